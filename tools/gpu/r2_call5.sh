#!/bin/bash
# round 2, call 5: ncu full capture of the C2 composites + launch list
cd "$GRAFT_REPO_ROOT"
ncu --set full --clock-control none --import-source on -k regex:composite -s 4 -c 2 -o gpurun_out/prof_r02a python tools/profile_frame.py --workload C2 --frames 3 > gpurun_out/ncu_r02a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02a.csv python tools/profile_frame.py --workload C2 --frames 3 > gpurun_out/launches_r02a.log 2>&1
tail -3 gpurun_out/ncu_r02a.log; ls -la gpurun_out/prof_r02a.ncu-rep
