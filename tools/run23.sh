#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_inference_api_gpu.py -x -q -k "legacy" 2>&1 | tail -30
