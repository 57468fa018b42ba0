#!/bin/bash
# round 2: the driver's scaling command lines at N = 8 (and the reference arm under torchrun)
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 8 --steps 5 --warmup 1 > gpurun_out/r2r_ref_n8.log 2>&1; echo "ref n8 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2r_n8.log 2>&1; echo "n8 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2r_ref_n8.log','gpurun_out/r2r_n8.log'):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True
            print(f, 'value %.2f'%d['value'], 'n', d['n_gpus'], 'e2e', d.get('e2e') and round(d['e2e']['value'],1), 'merged', d.get('e2e_merged') and round(d['e2e_merged']['value'],1), 'strong', d.get('strong_scaling') and round(d['strong_scaling']['value'],1), 'coll', d.get('collective'))
    if not ok: print(f, open(f).read()[-1500:])
PY
