#!/bin/bash
cd /root/repo
timeout 600 python tools/prof_unet.py 2>&1 | grep -v "initialize\|Warning" | tail -36
