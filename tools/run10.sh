#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -q -m gpu -s -p no:cacheprovider -k "unet or cascade" > gpurun_out/t_train_unet.log 2>&1; echo "unet train rc=$?"; grep -E "unet|loss |passed|failed|Error|error|assert" gpurun_out/t_train_unet.log | tail -40
timeout 600 python -m pytest tests/test_backward_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "norm_bwd" > gpurun_out/t_b.log 2>&1; echo "norm_bwd rc=$?"; tail -2 gpurun_out/t_b.log
