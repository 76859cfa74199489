#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -x -k "norm or stats" > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?"; tail -3 gpurun_out/t_kern.log
timeout 900 python -m pytest tests/test_resnet_gpu.py tests/test_unet_d_gpu.py -q -m gpu -s -p no:cacheprovider > gpurun_out/t_nets.log 2>&1; echo "nets rc=$?"; grep -E "max\|d\||passed|failed|Error" gpurun_out/t_nets.log | tail -30
for cfg in "4 1" "8 1" "8 2" "8 3" "16 2"; do set -- $cfg
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --micro-batch $1 --streams $2 > gpurun_out/bench_mb$1_s$2.json 2> gpurun_out/bench.err; echo "bench mb=$1 streams=$2 rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_mb$1_s$2.json"))
print("  value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],1), "host_ms", round(d["config"]["host_enqueue_ms_per_step"],1), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), d["clocks"])
PY
done
tail -3 gpurun_out/bench.err
