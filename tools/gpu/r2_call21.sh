#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "five" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2s_pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-single > gpurun_out/r2s_n1.log 2>&1; echo "n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-single > gpurun_out/r2s_n2.log 2>&1; echo "n2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-single --no-graph-collectives > gpurun_out/r2s_n2_nogc.log 2>&1; echo "n2 nogc rc=$?"
for f in gpurun_out/r2s_n1.log gpurun_out/r2s_n2.log gpurun_out/r2s_n2_nogc.log; do python - $f <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); ok=True
        print(sys.argv[1], 'value %.1f'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'strong', d['strong_scaling'] and round(d['strong_scaling']['value'],1), 'coll', d['collective'] and (round(d['collective']['ms_per_step'],4), d['collective']['in_step_graph']))
if not ok: print(open(sys.argv[1]).read()[-1500:])
PY
done
grep -i "capturing the collectives failed" gpurun_out/r2s_n2.log | head -2
