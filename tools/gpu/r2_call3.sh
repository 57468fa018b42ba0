#!/bin/bash
# round 2, call 3: light/heavy forward split, tile skipping, bit-exact centres: parity + A/B matrix
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2c_parity.jsonl 2>/dev/null
grep -h "passed\|failed\|^FAILED" gpurun_out/r2c_pytest.log | tail -8
B="--steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-single --no-e2e"
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B $ARGS > gpurun_out/r2c_$name.log 2>&1; python - gpurun_out/r2c_$name.log $name <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; ok=True
        pk=r['per_kernel_ms']
        print(f"{sys.argv[2]:28s} value {d['value']:8.1f}  fwd {pk.get('composite_fwd',0)*1e3:6.1f} bwd {pk.get('composite_bwd',0)*1e3:6.1f} us  frac {r['frac']:.4f} ({r['kernel']})  sum/frame {r.get('sum_kernel_ms_per_training_frame', r.get('sum_kernel_ms_per_frame',0))*1e3:7.1f} us")
if not ok: print(sys.argv[2], 'FAILED'); print(open(sys.argv[1]).read()[-600:])
PY
}
ARGS="--workload C2 --pattern single"
run c2_default X=1
run c2_heavy_off B2R_HEAVY=off
run c2_heavy_1024 B2R_HEAVY=1024
run c2_unsegmented B2R_SEGMENTED=0
run c2_unseg_heavyoff B2R_SEGMENTED=0 B2R_HEAVY=off
ARGS="--workload C4 --engine merged"
run c4_merged X=1
run c4_merged_noskip B2R_SKIP_TILES=0
run c4_merged_heavy_off B2R_HEAVY=off
run c4_merged_heavy_1024 B2R_HEAVY=1024
ARGS="--workload C4 --engine separate"
run c4_separate X=1
run c4_separate_v3like B2R_SEGMENTED=0 B2R_HEAVY=off
