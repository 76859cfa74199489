#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?"; tail -3 gpurun_out/t_train.log
timeout 600 python bench.py --workload train --steps 2 --warmup 2 2> gpurun_out/b.err | grep '^{' | tee gpurun_out/bench_train.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train:', round(d['value'],2), 'tiles/s', round(d['ms_per_step'],1),'ms/step host', round(d['config']['host_enqueue_ms_per_step'],1), 'launches', d['gpu_launches'])"
