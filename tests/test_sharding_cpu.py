"""World-size-2 gloo tests (CPU) of the N > 1 host logic: tile sharding, result gather in original order, the
max-over-ranks timing reduction, and `bench.py --impl reference` behaviour of non-zero ranks."""
import os
import socket
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from deepliif_b200 import sharding
    tiles = np.arange(n_items * 4 * 4 * 3, dtype=np.uint8).reshape(n_items, 4, 4, 3)
    mine = sharding.shard(tiles, rank, world)
    assert [int(t[0, 0, 0]) for t in mine] == [int(tiles[i][0, 0, 0]) for i in sharding.shard_indices(n_items, rank, world)]
    processed = 255 - mine                               # stand-in for the per-rank pipeline
    full = sharding.gather_to_rank0(processed, n_items)
    slow = sharding.max_over_ranks(10.0 + rank)
    if rank == 0:
        q.put((np.array_equal(full, 255 - tiles), slow))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_gather_world2_gloo():
    for n_items in (7, 8, 1):
        ctx = mp.get_context("spawn")
        q = ctx.SimpleQueue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        ok, slow = q.get()
        assert ok and slow == 11.0


def test_shard_indices_cover_everything_once():
    from deepliif_b200 import sharding
    for n in (0, 1, 5, 83, 102):
        for w in (1, 2, 4, 8):
            idx = sorted(i for r in range(w) for i in sharding.shard_indices(n, r, w))
            assert idx == list(range(n))


def test_reference_arm_nonzero_rank_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
