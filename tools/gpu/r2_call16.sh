#!/bin/bash
cd "$GRAFT_REPO_ROOT"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-single --no-e2e"
for mc in 8 16 32; do for l in 3 4 6; do
  CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 600 python bench.py $B --lanes $l > gpurun_out/r2o_c4_mc${mc}_l$l.log 2>&1
  python - $mc $l <<'PY'
import json,sys
for l in open(f'gpurun_out/r2o_c4_mc{sys.argv[1]}_l{sys.argv[2]}.log'):
    if l.startswith('{'):
        d=json.loads(l); print('max_connections', sys.argv[1], 'lanes', sys.argv[2], 'value %.1f'%d['value'])
PY
done; done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 python bench.py $B --workload C2 --pattern single > gpurun_out/r2o_c2_mc32.log 2>&1; grep -o '"value": [0-9.]*' gpurun_out/r2o_c2_mc32.log | head -1
