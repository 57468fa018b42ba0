#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "training_frame or five_render" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2l_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single > gpurun_out/r2l_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
ok=False
for l in open('gpurun_out/r2l_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); ok=True
        print('value %.1f e2e %.1f e2e_merged %s eager %.1f'%(d['value'], d['e2e']['value'], d['e2e_merged'] and round(d['e2e_merged']['value'],1), d['e2e_eager']['value']))
if not ok: print(open('gpurun_out/r2l_bench.log').read()[-1500:])
PY
grep -i "e2e_merged leg failed" gpurun_out/r2l_bench.log
