#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "conv_tc" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"; tail -6 gpurun_out/t_tc.log
timeout 900 python -m pytest tests/test_resnet_gpu.py tests/test_unet_d_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_nets.log 2>&1; echo "nets rc=$?"; tail -3 gpurun_out/t_nets.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_vs.json 2> gpurun_out/bench.err; echo "bench rc=$?"; grep '^{' gpurun_out/bench_vs.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))"; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 300 --csv --log-file gpurun_out/launches_vs.csv python bench.py --steps 1 --warmup 1 --batch 8 --micro-batch 8 --streams 1 --no-cpu-baseline --no-roofline-events > gpurun_out/ncu_launch.log 2>&1; echo "ncu rc=$?"
