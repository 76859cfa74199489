#!/bin/bash
cd /root/repo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_mt_kernel -s 2 -c 1 -o gpurun_out/r01_wgrad_mt python tools/wgrad_one.py > gpurun_out/ncu_wgrad.log 2>&1
tail -3 gpurun_out/ncu_wgrad.log
timeout 600 python tools/prof_train.py 2>&1 | grep -v "initialize\|Initializing\|created\|warn" | tail -26 > gpurun_out/train_breakdown.txt; cat gpurun_out/train_breakdown.txt | head -14
