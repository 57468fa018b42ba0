#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single --no-e2e > gpurun_out/r2p_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
ok=False
for l in open('gpurun_out/r2p_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); ok=True
        print('value %.1f'%d['value']); 
        for k,v in d['roofline']['per_view'].items(): print(k, {w:(round(x['ms']*1e3,1), int(x['consumed']), round(x['frac'],4)) for w,x in v.items()})
if not ok: print(open('gpurun_out/r2p_bench.log').read()[-1500:])
PY
