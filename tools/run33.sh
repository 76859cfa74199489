#!/bin/bash
cd /root/repo
timeout 600 python bench.py --workload train --graph --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/tr.err | grep "^{" | tee gpurun_out/bench_train_graph.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['config'].get('host_enqueue_ms_per_step'))"
tail -3 gpurun_out/tr.err | cut -c1-300
