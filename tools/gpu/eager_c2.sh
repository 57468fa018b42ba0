#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/eager_profile.py --workload C2 --single --frames 200 > gpurun_out/$1_eager_c2.txt 2>&1; head -1 gpurun_out/$1_eager_c2.txt
timeout 300 python tools/eager_profile.py --workload C2 --single --same-camera --frames 200 > gpurun_out/$1_eager_c2_samecam.txt 2>&1; head -1 gpurun_out/$1_eager_c2_samecam.txt
B2R_COMPILED_BINDING=0 timeout 300 python tools/eager_profile.py --workload C2 --single --same-camera --frames 200 2>&1 | head -1
