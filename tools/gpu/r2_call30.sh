#!/bin/bash
# ncu source-level capture of the forward composite (packed build) on the five views of a C4 frame
cd "$GRAFT_REPO_ROOT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:composite_fwd -s 15 -c 5 -o gpurun_out/prof_c30_fwd python tools/five_breakdown.py > gpurun_out/ncu_c30.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_c30.log; ls -la gpurun_out/prof_c30_fwd.ncu-rep
