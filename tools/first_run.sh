#!/bin/bash
# first GPU bring-up: every stage in its own process so a trap in one does not poison the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -m pytest tests/test_pixel_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_pixel.log 2>&1; echo "pixel rc=$?"
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "not conv_tc" -p no:cacheprovider > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?"
for c in k1s1_64_64 k3s1_64_64_w16 k3s1_256_256_w128 k3s2_64_128 ct3s2_256_128; do
  timeout 120 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_tc and $c and bf16x3" -p no:cacheprovider > gpurun_out/t_tc_$c.log 2>&1; echo "tc $c rc=$?"
done
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv_tc" -p no:cacheprovider > gpurun_out/t_tc_all.log 2>&1; echo "tc all rc=$?"
timeout 600 python -m pytest tests/test_resnet_gpu.py -q -m gpu -k "direct" -s -p no:cacheprovider > gpurun_out/t_resnet_direct.log 2>&1; echo "resnet direct rc=$?"
timeout 900 python -m pytest tests/test_resnet_gpu.py -q -m gpu -k "not direct" -s -p no:cacheprovider > gpurun_out/t_resnet_tc.log 2>&1; echo "resnet tc rc=$?"
tail -n 5 gpurun_out/t_*.log
