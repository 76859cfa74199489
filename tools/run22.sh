#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_training_gpu.py tests/test_inference_api_gpu.py -x -q 2>&1 | tail -8
