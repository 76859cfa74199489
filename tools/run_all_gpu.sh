#!/bin/bash
mkdir -p gpurun_out
timeout 3000 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --durations=15 > gpurun_out/t_all.log 2>&1; echo "all gpu tests rc=$?"; tail -30 gpurun_out/t_all.log | grep -v "^$"
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
