#!/bin/bash
# round 2, final build: the round-end sequence the driver runs (tests, smoke, reference arm, bench) + the ncu evidence
# (launch list of the bench command, full captures of the composites on the C4 five-render frame, C2 and C3)
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2y_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2y_parity.jsonl
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_ref.log 2>&1; echo "ref rc=$?"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_bench.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
ref=None
for l in open('gpurun_out/r2y_ref.log'):
    if l.startswith('{'): ref=json.loads(l)
for l in open('gpurun_out/r2y_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value %.1f  e2e %.1f  e2e_merged %.1f  eager %.1f  single(C2) %.1f  cpu %.3f  ref %.3f'%(d['value'], d['e2e']['value'], d['e2e_merged']['value'], d['e2e_eager']['value'], d['single_render']['value'], d['cpu_baseline']['value'], ref['value']))
        print('roofline', r['kernel'], 'frac %.4f'%r['frac'], 'ms %.4f'%r['kernel_ms_avg'], 'traffic', r['traffic'], 'other', r['other_composite'])
        print('per kernel', r['per_kernel_ms']); print('clocks', d['clocks']); print('launches', d['gpu_launches'], 'ms/step', d['ms_per_step'])
        print('single', json.dumps(d['single_render']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02b_c4five.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-single --no-eager --no-graph > gpurun_out/launches_r02b_c4five.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:composite -s 30 -c 10 -o gpurun_out/prof_r02b_c4five python tools/five_breakdown.py > gpurun_out/ncu_r02b_c4five.log 2>&1
for wl in C2 C3; do
  ncu --set full --clock-control none -k regex:composite -s 4 -c 2 -o gpurun_out/prof_r02b_$wl python tools/profile_frame.py --workload $wl --frames 3 > gpurun_out/ncu_r02b_$wl.log 2>&1
done
python tools/five_breakdown.py > gpurun_out/r2y_five_breakdown.txt 2>&1
ls -la gpurun_out/*r02b*.ncu-rep
