#!/bin/bash
# ncu --set full (source-level) of the ten composite launches of one C4 training frame (five_breakdown: 3 warm-up frames
# = 30 composite launches, then the profiled frame).  usage: ncu_composites.sh <tag>
cd "$GRAFT_REPO_ROOT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:composite -s 30 -c 10 -o gpurun_out/prof_$1 python tools/five_breakdown.py > gpurun_out/ncu_$1.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$1.log; ls -la gpurun_out/prof_$1.ncu-rep
