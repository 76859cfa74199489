#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/t_train.log 2>&1; echo "train rc=$?"; grep -E "worst|loss |cosine|passed|failed|Error|assert" gpurun_out/t_train.log | tail -40
timeout 900 python bench.py --workload train --steps 2 --warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; tail -1 gpurun_out/bench_train.json | cut -c1-900; tail -5 gpurun_out/bench_train.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
