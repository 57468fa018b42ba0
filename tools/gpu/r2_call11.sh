#!/bin/bash
# round 2, call 11: split forward of long lists: parity + A/B
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest(split 1024) rc=$?"; tail -3 gpurun_out/r2j_pytest.log
B2R_SPLIT=512 timeout 1800 python -m pytest tests -m gpu -q -k "fullsize or long or five" > gpurun_out/r2j_pytest512.log 2>&1; echo "pytest(split 512) rc=$?"; tail -3 gpurun_out/r2j_pytest512.log
B="--steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-single --no-e2e"
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B $ARGS > gpurun_out/r2j_$name.log 2>&1; python - gpurun_out/r2j_$name.log $name <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; ok=True
        pk=r['per_kernel_ms']
        print(f"{sys.argv[2]:24s} value {d['value']:8.1f} ms/step {d['ms_per_step']:.4f} fwd {pk.get('composite_fwd',0)*1e3:6.1f} bwd {pk.get('composite_bwd',0)*1e3:6.1f} us frac {r['frac']:.4f} ({r['kernel'][10:]})")
if not ok: print(sys.argv[2], 'FAILED'); print(open(sys.argv[1]).read()[-800:])
PY
}
for sp in off 2048 1024 512; do
  ARGS="--workload C2 --pattern single --lanes 1 --frames 1"; run c2_solo_$sp B2R_SPLIT=$sp
  ARGS="--workload C2 --pattern single"; run c2_l4_$sp B2R_SPLIT=$sp
  ARGS="--workload C4"; run c4_merged_$sp B2R_SPLIT=$sp
  ARGS="--workload C4 --lanes 1 --frames 1"; run c4_solo_$sp B2R_SPLIT=$sp
  ARGS="--workload C5 --pattern single"; run c5_l4_$sp B2R_SPLIT=$sp
done
B2R_SPLIT=1024 python tools/five_breakdown.py > gpurun_out/r2j_breakdown_1024.txt 2>&1; grep "forward\|backward" gpurun_out/r2j_breakdown_1024.txt
