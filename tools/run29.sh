#!/bin/bash
cd /root/repo
timeout 600 python tools/prof_host.py 2>&1 | grep -v "initialize\|Initializing\|created\|warn" | cut -c1-150 | tail -90
