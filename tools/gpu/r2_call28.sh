#!/bin/bash
# compiled torch binding of the eager call: full GPU suite (both routes), eager profile, bench with the eager legs
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c28_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c28_pytest.log
timeout 300 python tools/eager_profile.py --frames 40 > gpurun_out/c28_eager_five.txt 2>&1; head -1 gpurun_out/c28_eager_five.txt
B2R_COMPILED_BINDING=0 timeout 300 python tools/eager_profile.py --frames 40 2>&1 | head -1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c28_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/c28_bench.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value %.1f  e2e %.1f  e2e_merged %.1f'%(d['value'], d['e2e']['value'], d['e2e_merged']['value']), 'eager', json.dumps(d['e2e_eager'])[:300])
PY
