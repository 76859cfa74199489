#!/bin/bash
cd /root/repo
timeout 600 python bench.py --workload postprocess --steps 3 --warmup 3 2> gpurun_out/pp.err | grep "^{" | tee gpurun_out/bench_postprocess.json
tail -5 gpurun_out/pp.err
timeout 600 python -m pytest tests/test_inference_api_gpu.py -x -q 2>&1 | tail -5
