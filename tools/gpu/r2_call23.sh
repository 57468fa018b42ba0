#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "training_frame" > gpurun_out/c23_tests.txt 2>&1; tail -5 gpurun_out/c23_tests.txt
timeout 300 python tools/eager_profile.py --frames 60 --fused --graph > gpurun_out/eager_profile_fused_graph.txt 2>&1
head -2 gpurun_out/eager_profile_fused_graph.txt
