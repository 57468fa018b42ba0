#!/bin/bash
# round 2, call 9 (2 GPUs): the driver's own command lines, N = 1 and N = 2
cd "$GRAFT_REPO_ROOT"
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2h_n1.log 2>&1; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2h_n2.log 2>&1; echo "n2 rc=$?"
for f in gpurun_out/r2h_n1.log gpurun_out/r2h_n2.log; do python - $f <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); ok=True
        print(sys.argv[1], 'value %.1f'%d['value'], 'e2e', d['e2e'] and round(d['e2e']['value'],1), 'eager', d['e2e_eager'] and round(d['e2e_eager']['value'],1), 'strong', d['strong_scaling'] and round(d['strong_scaling']['value'],1), 'coll', d['collective'], 'single', d['single_render'] and round(d['single_render']['value'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],3))
if not ok: print(open(sys.argv[1]).read()[-1500:])
PY
done
