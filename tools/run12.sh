#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 4" "4 3" "4 6" "8 3" "8 4"; do set -- $cfg
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-roofline-events --micro-batch $1 --streams $2 > gpurun_out/sw_mb$1_s$2.json 2> gpurun_out/bench.err; echo "bench mb=$1 streams=$2 rc=$?"
  grep '^{' gpurun_out/sw_mb$1_s$2.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'host_ms', round(d['config']['host_enqueue_ms_per_step'],1), d['clocks']['sm_mhz'])"
done
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench.err; echo "default rc=$?"; grep '^{' gpurun_out/bench_default.json | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['roofline'])"
