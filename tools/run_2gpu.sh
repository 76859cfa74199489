#!/bin/bash
# 2-GPU tier: NCCL tests, tile-sharded inference bench, data-parallel training bench (launched as the driver does).
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/s2.err | grep "^{" | tee gpurun_out/scale_2gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 4 --warmup 3 --no-cpu-baseline 2>> gpurun_out/s2.err | grep "^{" | tee gpurun_out/train_2gpu.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step')})"
tail -2 gpurun_out/s2.err | cut -c1-200
