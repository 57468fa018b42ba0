#!/bin/bash
cd "$GRAFT_REPO_ROOT"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-single --no-e2e"
for l in 2 3 4 5 6 8; do
  timeout 600 python bench.py $B --lanes $l > gpurun_out/r2g_c4_l$l.log 2>&1
  python - $l <<'PY'
import json,sys
for l in open(f'gpurun_out/r2g_c4_l{sys.argv[1]}.log'):
    if l.startswith('{'):
        d=json.loads(l); print('lanes', sys.argv[1], 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'])
PY
done
