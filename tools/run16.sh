#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload train --steps 2 --warmup 2 2> gpurun_out/b.err | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train:', round(d['value'],2), 'tiles/s', round(d['ms_per_step'],1),'ms/step host', round(d['config']['host_enqueue_ms_per_step'],1), 'launches', d['gpu_launches'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 2600 --csv --log-file gpurun_out/launches_train.csv python bench.py --workload train --steps 1 --warmup 1 --batch 2 > gpurun_out/ncu_train.log 2>&1; echo "ncu rc=$?"
