#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_training_gpu.py tests/test_backward_kernels_gpu.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/tr.err | grep "^{" | tee gpurun_out/bench_train.json | cut -c1-420
timeout 600 python bench.py --workload train --topology default --steps 5 --warmup 3 --no-cpu-baseline 2>> gpurun_out/tr.err | grep "^{" | tee gpurun_out/bench_train_default.json | cut -c1-420
