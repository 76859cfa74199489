#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_d_gpu.py -q -m gpu -s -p no:cacheprovider -k "single_pass" 2>&1 | grep -E "unet256|passed|failed|Error" | tail -5
timeout 600 python bench.py --workload unet256 --steps 5 --warmup 3 2> gpurun_out/b.err | grep '^{' | cut -c1-700; tail -3 gpurun_out/b.err
timeout 600 python bench.py --workload unet256 --precision bf16x3 --steps 5 --warmup 3 2> gpurun_out/b.err | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bf16x3 unet256:', round(d['value'],1), 'tiles/s')" 
timeout 600 python bench.py --workload train --steps 2 --warmup 2 2> gpurun_out/b.err | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train:', round(d['value'],2), 'tiles/s', round(d['ms_per_step'],1),'ms/step')"
