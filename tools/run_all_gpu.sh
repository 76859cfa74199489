#!/bin/bash
# The 1-GPU tier as the driver runs it: GPU tests, smoke, default bench (own arm) + reference arm.
cd /root/repo; mkdir -p gpurun_out
timeout 3000 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=8 > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"; tail -14 gpurun_out/t_all.log | grep -v "^$"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/bench.err | grep "^{" | tee gpurun_out/bench_1gpu.json | cut -c1-1500
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 2>> gpurun_out/bench.err | grep "^{" | tee gpurun_out/bench_ref.json | cut -c1-600
