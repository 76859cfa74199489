#!/bin/bash
cd /root/repo
timeout 600 python tools/prof_wgrad.py 2>&1 | grep -v Warning | tail -60
