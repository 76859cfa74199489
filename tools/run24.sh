#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_training_gpu.py -x -q -s -k "resnet or padding_backward" 2>&1 | grep -v Warning | tail -25
