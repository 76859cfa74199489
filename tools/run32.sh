#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_training_gpu.py -x -q -k "cuda_graph" 2>&1 | grep -v "Warning\|initialize\|Initializing\|created" | tail -5
