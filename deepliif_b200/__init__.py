"""deepliif_b200 — B200-native (sm_100a) implementation of the DeepLIIF tile-parallel cGAN hot path.

Host side is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic of the path runs in
hand-written CUDA kernels behind the C ABI declared in include/deepliif_b200.h (libdeepliif_b200.so).
There is no CPU or library fallback: importing `deepliif_b200.ops` without the built library raises.
"""
__version__ = "0.1.0"
