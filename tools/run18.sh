#!/bin/bash
mkdir -p gpurun_out
for cfg in "4 3 128" "2 3 128" "4 3 0" "8 3 128"; do set -- $cfg
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-roofline-events --micro-batch $1 --streams $2 --trunk-n-tile $3 2> gpurun_out/bench.err | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mb=$1 s=$2 ntile=$3: value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done
