"""torchrun target of `deepliif trainlaunch` (twin of the reference's deepliif/scripts/train.py): forwards its
argv to the `train` command inside each rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepliif_b200.cli import cli  # noqa: E402

if __name__ == "__main__":
    cli(["train", *sys.argv[1:]])
