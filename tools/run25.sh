#!/bin/bash
cd /root/repo
timeout 600 python tools/adhoc_reflect.py 2>&1 | grep -v "worst max-abs" | tail -8
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
