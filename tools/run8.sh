#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload train --steps 2 --warmup 2 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench train rc=$?"; tail -1 gpurun_out/bench_train.json | cut -c1-1200; tail -3 gpurun_out/bench_train.err
timeout 600 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider -k gradients > gpurun_out/t_train.log 2>&1; echo "train grads rc=$?"; tail -2 gpurun_out/t_train.log
