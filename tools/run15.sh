#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/racecheck.log
