#!/bin/bash
# round 2, call 1: new full-size parity tests + new bench on the round-1 kernels
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
nproc >> gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --engine separate > gpurun_out/r2a_bench_c4.log 2>&1; echo "bench rc=$?" >> gpurun_out/r2a_bench_c4.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.log 2>&1
tail -5 gpurun_out/r2a_pytest.log; tail -c 1500 gpurun_out/r2a_bench_c4.log
