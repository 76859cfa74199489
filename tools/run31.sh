#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_training_gpu.py -x -q -k "cuda_graph or dropout" 2>&1 | grep -v Warning | tail -25
timeout 600 python bench.py --workload train --topology default --graph --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/tr.err | grep "^{" | tee gpurun_out/bench_train_default_graph.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['config'].get('host_enqueue_ms_per_step'))"
tail -5 gpurun_out/tr.err
