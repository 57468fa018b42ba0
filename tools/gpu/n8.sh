#!/bin/bash
# two GPUs: the driver's launch line, weak + strong + collectives
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-eager > gpurun_out/$1_n8.log 2>&1; echo "rc=$?"
python - gpurun_out/$1_n8.log <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('N=8 value %.1f ms/step %.3f e2e %.1f merged %.1f'%(d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('e2e_merged') or {}).get('value',0)))
        print('strong', json.dumps(d.get('strong_scaling'))[:400]); print('collective', json.dumps(d.get('collective'))[:400])
PY
tail -2 gpurun_out/$1_n8.log | cut -c1-200
