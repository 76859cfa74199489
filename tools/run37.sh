#!/bin/bash
cd /root/repo
timeout 600 python bench.py --workload cascade --steps 3 --warmup 3 2> gpurun_out/c.err | grep "^{" | tee gpurun_out/bench_cascade.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})"
timeout 600 python bench.py --workload train --topology default --graph --steps 5 --warmup 3 --no-cpu-baseline 2>> gpurun_out/c.err | grep "^{" | tee gpurun_out/bench_train_default_graph.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})"
timeout 600 python bench.py --workload train --topology default --steps 5 --warmup 3 --no-cpu-baseline 2>> gpurun_out/c.err | grep "^{" | tee gpurun_out/bench_train_default.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})"
timeout 600 python bench.py --workload unet256 --steps 5 --warmup 3 2>> gpurun_out/c.err | grep "^{" | tee gpurun_out/bench_unet256.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})"
tail -3 gpurun_out/c.err | cut -c1-200
