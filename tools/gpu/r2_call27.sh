#!/bin/bash
# adaptive private/cooperative tile walks + variants (register caps, walk cost constants)
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c27_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c27_pytest.log
for v in default coop w40 w120 pb3 pj5; do
  if [ $v = default ]; then unset B2R_LIB; else export B2R_LIB=$PWD/gpurun_variants/lib_$v.so; fi
  echo "=== $v"
  timeout 300 python tools/five_breakdown.py 2>&1 | grep "project + bin\|backward projection" 
  timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-single --no-eager 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.1f'%d['value'], d['roofline']['per_kernel_ms'])"
done
