#!/bin/bash
cd /root/repo
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/tr.err | grep "^{" | tee gpurun_out/bench_train.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, {k:v for k,v in d.items() if 'host' in k})"
timeout 600 python bench.py --workload train --topology default --steps 5 --warmup 3 --no-cpu-baseline 2>> gpurun_out/tr.err | grep "^{" | tee gpurun_out/bench_train_default.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, {k:v for k,v in d.items() if 'host' in k})"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>> gpurun_out/tr.err | grep "^{" | tee gpurun_out/bench_inf.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','e2e')})"
