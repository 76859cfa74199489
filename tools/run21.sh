#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_cells_gpu.py -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload postprocess --steps 5 --warmup 3 2> gpurun_out/pp.err | grep "^{" | tee gpurun_out/bench_postprocess.json
tail -3 gpurun_out/pp.err
