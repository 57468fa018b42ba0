#!/bin/bash
# round 2, call 12: posed-gradient test, compute-sanitizer on the new kernels
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_fused_skinning.py -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2k_pytest.log
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_r02.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "five_render_plan or long_list_sort_is_bit_exact and 3000 or T1 or accumulate" > gpurun_out/r2k_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2k_memcheck_pytest.log; grep -c "ERROR SUMMARY" gpurun_out/sanitizer_memcheck_r02.log; tail -2 gpurun_out/sanitizer_memcheck_r02.log
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_r02.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "five_render_plan and merged or long_list_sort_is_bit_exact and 600" > gpurun_out/r2k_racecheck_pytest.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r2k_racecheck_pytest.log; tail -2 gpurun_out/sanitizer_racecheck_r02.log
