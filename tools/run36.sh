#!/bin/bash
cd /root/repo
timeout 600 python tools/prof_unet.py 2>&1 | grep -v "initialize\|Warning" | head -24
timeout 900 python -m pytest tests/test_unet_d_gpu.py tests/test_kernels_gpu.py -x -q 2>&1 | tail -3
