#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 450 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --batch 4 --micro-batch 4 --no-cpu-baseline --no-roofline-events > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 12 -c 2 -o gpurun_out/prof_conv_tc python bench.py --steps 1 --warmup 1 --batch 4 --micro-batch 4 --no-cpu-baseline --no-roofline-events > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
