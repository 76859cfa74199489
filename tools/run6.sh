#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py -q -m gpu -s -p no:cacheprovider -k "gradients" > gpurun_out/t_train1.log 2>&1; echo "train grads rc=$?"; grep -E "worst|passed|failed|Error|error|assert" gpurun_out/t_train1.log | tail -30
timeout 1200 python -m pytest tests/test_training_gpu.py -q -m gpu -s -p no:cacheprovider -k "optimisation" > gpurun_out/t_train2.log 2>&1; echo "train step rc=$?"; grep -E "loss |weights off|passed|failed|Error|error|assert" gpurun_out/t_train2.log | tail -40
timeout 600 python -m pytest tests/test_unet_d_gpu.py -q -m gpu -s -p no:cacheprovider -k "discriminator" > gpurun_out/t_d.log 2>&1; echo "D rc=$?"; grep -E "max\|d\||passed|failed" gpurun_out/t_d.log | tail
