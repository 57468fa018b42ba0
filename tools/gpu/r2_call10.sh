#!/bin/bash
# round 2, call 10: full GPU suite after the camera change, eager number, ncu evidence for profiles/
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2i_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2i_parity.jsonl
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single > gpurun_out/r2i_bench.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2i_bench.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value %.1f e2e %.1f eager %.1f host_ms/render %.3f'%(d['value'], d['e2e']['value'], d['e2e_eager']['value'], d['e2e_eager']['host_ms_per_render']))
PY
# launch list of the bench command (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_c4five.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-single --no-eager --no-graph > gpurun_out/launches_r02_c4five.log 2>&1
# full captures: C4 five-render views (serial), C2 / C3 / C5 single frames
ncu --set full --clock-control none --import-source on -k regex:composite -s 30 -c 10 -o gpurun_out/prof_r02_c4five python tools/five_breakdown.py > gpurun_out/ncu_r02_c4five.log 2>&1
for wl in C2 C3 C5; do
  ncu --set full --clock-control none -k regex:composite -s 4 -c 2 -o gpurun_out/prof_r02_$wl python tools/profile_frame.py --workload $wl --frames 3 > gpurun_out/ncu_r02_$wl.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
