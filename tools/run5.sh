#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_kernels_gpu.py -q -m gpu -s -p no:cacheprovider -k "norm_bwd or head_bwd" > gpurun_out/t_bwd1.log 2>&1; echo "bwd norm rc=$?"; tail -5 gpurun_out/t_bwd1.log
timeout 300 python -m pytest tests/test_backward_kernels_gpu.py -q -m gpu -s -p no:cacheprovider -k "wgrad and k3s1_64_64" > gpurun_out/t_bwd2.log 2>&1; echo "bwd wgrad first rc=$?"; grep -E "max\|d\||passed|failed|Error|assert" gpurun_out/t_bwd2.log | tail
timeout 600 python -m pytest tests/test_backward_kernels_gpu.py -q -m gpu -s -p no:cacheprovider -k "wgrad" > gpurun_out/t_bwd3.log 2>&1; echo "bwd wgrad all rc=$?"; grep -E "max\|d\||passed|failed|Error|assert" gpurun_out/t_bwd3.log | tail -40
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "norm or stats" > gpurun_out/t_kern.log 2>&1; echo "kern rc=$?"; tail -3 gpurun_out/t_kern.log
