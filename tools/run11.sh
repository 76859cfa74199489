#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -q -m gpu -s -p no:cacheprovider -k "dropout or unet or resnet_generator" > gpurun_out/t_drop.log 2>&1; echo "dropout rc=$?"; grep -E "unet|resnet|passed|failed|Error|error|assert" gpurun_out/t_drop.log | tail -20
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_backward_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "norm" > gpurun_out/t_n.log 2>&1; echo "norm rc=$?"; tail -2 gpurun_out/t_n.log
