#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -q -x -k "five or c4" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2f_pytest.log
python tools/five_breakdown.py > gpurun_out/r2f_breakdown.txt 2>&1; cat gpurun_out/r2f_breakdown.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-single --no-e2e > gpurun_out/r2f_c4.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/r2f_c4.log'):
    if l.startswith('{'):
        d=json.loads(l); print('C4 merged value', d['value'], 'ms/step', d['ms_per_step'], d['roofline']['per_kernel_ms'])
PY
