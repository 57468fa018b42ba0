#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python bench.py --steps 20 --warmup 5 --no-single --no-cpu-baseline > gpurun_out/$1_e2e.log 2>&1; echo rc=$?
tail -3 gpurun_out/$1_e2e.log | cut -c1-300
