#!/bin/bash
# ncu full capture of the chain kernels (projection, scan, scatter, sort, merge, backward projection) on the C4 merged pass
cd "$GRAFT_REPO_ROOT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'project|scatter|sort|scan|merge' -s 12 -c 7 -o gpurun_out/prof_r02_chain python tools/five_breakdown.py > gpurun_out/ncu_r02_chain.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r02_chain.log; ls -la gpurun_out/prof_r02_chain.ncu-rep
