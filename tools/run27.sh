#!/bin/bash
cd /root/repo
timeout 600 python -m pytest tests/test_backward_kernels_gpu.py -x -q -s -k wgrad 2>&1 | grep -v Warning | grep "wgrad\|passed\|failed\|Error\|error" | tail -30
timeout 600 python tools/prof_wgrad.py 2>&1 | grep "== wgrad" -A 9
