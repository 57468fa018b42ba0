#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python tools/eager_profile.py --frames 40 > gpurun_out/eager_profile_five.txt 2>&1
python tools/eager_profile.py --frames 40 --fused > gpurun_out/eager_profile_fused.txt 2>&1
head -3 gpurun_out/eager_profile_five.txt; head -3 gpurun_out/eager_profile_fused.txt
