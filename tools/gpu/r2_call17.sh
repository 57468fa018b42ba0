#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python tools/five_timeline.py --lanes 3 > gpurun_out/five_timeline_l3.txt 2>&1; cat gpurun_out/five_timeline_l3.txt | tail -40
python tools/five_timeline.py --lanes 1 > gpurun_out/five_timeline_l1.txt 2>&1; head -1 gpurun_out/five_timeline_l1.txt; tail -3 gpurun_out/five_timeline_l1.txt
