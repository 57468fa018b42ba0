#!/bin/bash
# round 2: the round-end sequence the driver runs (tests, smoke, reference arm, bench), on the final build
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2x_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2x_parity.jsonl
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2x_ref.log 2>&1; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2x_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
ref=None
for l in open('gpurun_out/r2x_ref.log'):
    if l.startswith('{'): ref=json.loads(l)
for l in open('gpurun_out/r2x_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value %.1f  e2e %.1f  e2e_merged %.1f  eager %.1f  single(C2) %.1f  cpu %.3f  ref %.3f'%(d['value'], d['e2e']['value'], d['e2e_merged']['value'], d['e2e_eager']['value'], d['single_render']['value'], d['cpu_baseline']['value'], ref['value']))
        print('roofline', r['kernel'], 'frac %.4f'%r['frac'], 'ms %.4f'%r['kernel_ms_avg'], 'traffic', r['traffic'], 'other', r['other_composite'])
        print('per kernel', r['per_kernel_ms']); print('clocks', d['clocks']); print('launches', d['gpu_launches'], 'ms/step', d['ms_per_step'])
PY
