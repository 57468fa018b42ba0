#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python bench.py --workload C2 --pattern single --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$1_c2.log 2>&1; echo rc=$?
python - gpurun_out/$1_c2.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('value %.1f e2e %.1f'%(d['value'], d['e2e']['value']), 'eager', json.dumps(d.get('e2e_eager'))[:400])
PY
tail -3 gpurun_out/$1_c2.log | cut -c1-300
