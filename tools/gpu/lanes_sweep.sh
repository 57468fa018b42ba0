#!/bin/bash
# frames in flight re-tuned on the current build; full GPU suite first
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/$1_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/$1_pytest.log
for l in 2 3 4 5; do
  timeout 400 python bench.py --steps 20 --warmup 5 --lanes $l --no-e2e --no-eager --no-single --no-cpu-baseline > gpurun_out/$1_lanes$l.log 2>&1
  python - gpurun_out/$1_lanes$l.log $l <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('lanes', sys.argv[2], 'value %.1f ms/step %.3f'%(d['value'], d['ms_per_step']))
PY
done
