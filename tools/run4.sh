#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_inference_api_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/t_api.log 2>&1; echo "api rc=$?"; grep -E "mismatch|passed|failed|Error|error" gpurun_out/t_api.log | tail -20
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
# launch list (serialised, cold cache): shares only
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --batch 8 --micro-batch 8 --streams 1 --no-cpu-baseline --no-roofline-events > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# full capture of the ResNet-block conv (skip stem + 2 down convs of the first pass)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 5 -c 2 -o gpurun_out/prof_block_conv_r1 python bench.py --steps 1 --warmup 1 --batch 8 --micro-batch 8 --streams 1 --no-cpu-baseline --no-roofline-events > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
