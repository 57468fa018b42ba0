#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for l in 4 6 8; do
  timeout 600 python bench.py --steps 20 --warmup 5 --e2e-lanes $l --no-eager --no-single --no-cpu-baseline > gpurun_out/$1_e2e_l$l.log 2>&1
  python - gpurun_out/$1_e2e_l$l.log $l <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('e2e lanes', sys.argv[2], 'value %.1f e2e %.1f merged %.1f'%(d['value'], d['e2e']['value'], d['e2e_merged']['value']))
PY
done
