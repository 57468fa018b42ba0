#!/bin/bash
# warp-aggregated counter adds (projection count, scatter) + peer masks cached between the radix phases
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c25_pytest.log
timeout 300 python tools/five_breakdown.py > gpurun_out/c25_breakdown.txt 2>&1; head -20 gpurun_out/c25_breakdown.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-eager > gpurun_out/c25_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/c25_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value %.1f  e2e %.1f  e2e_merged %.1f single(C2) %s'%(d['value'], d['e2e']['value'], d['e2e_merged']['value'], json.dumps(d['single_render'])[:200]))
        print('per kernel', r['per_kernel_ms']); print('clocks', d['clocks'])
PY
