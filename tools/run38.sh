#!/bin/bash
cd /root/repo
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py > gpurun_out/memcheck.txt 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/memcheck.txt; tail -4 gpurun_out/memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize.py > gpurun_out/racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck.txt
