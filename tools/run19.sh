#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_cells_gpu.py -x -q 2>&1 | tail -25
