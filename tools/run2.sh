#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_pixel_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?"; tail -4 gpurun_out/t_kern.log
timeout 900 python -m pytest tests/test_resnet_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/t_resnet.log 2>&1; echo "resnet rc=$?"; tail -16 gpurun_out/t_resnet.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 450 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch 4 --micro-batch 4 --no-cpu-baseline --no-roofline-events > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
