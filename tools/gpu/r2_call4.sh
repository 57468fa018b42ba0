#!/bin/bash
# round 2, call 4: parity after the decision-path fix, frames in flight for the five-render pattern
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2d_parity.jsonl 2>/dev/null
grep -h "passed\|failed\|^FAILED" gpurun_out/r2d_pytest.log | tail -8
B="--steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-single --no-e2e"
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B $ARGS > gpurun_out/r2d_$name.log 2>&1; python - gpurun_out/r2d_$name.log $name <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; ok=True
        pk=r['per_kernel_ms']
        print(f"{sys.argv[2]:28s} value {d['value']:8.1f}  fwd {pk.get('composite_fwd',0)*1e3:6.1f} bwd {pk.get('composite_bwd',0)*1e3:6.1f} us  frac {r['frac']:.4f} ({r['kernel']})  sum/frame {r.get('sum_kernel_ms_per_training_frame', r.get('sum_kernel_ms_per_frame',0))*1e3:7.1f} us")
if not ok: print(sys.argv[2], 'FAILED'); print(open(sys.argv[1]).read()[-800:])
PY
}
ARGS="--workload C4 --engine merged --lanes 1"; run c4_merged_l1 X=1
ARGS="--workload C4 --engine merged --lanes 2"; run c4_merged_l2 X=1
ARGS="--workload C4 --engine merged --lanes 3"; run c4_merged_l3 X=1
ARGS="--workload C4 --engine merged --lanes 4"; run c4_merged_l4 X=1
ARGS="--workload C4 --engine separate --lanes 2"; run c4_separate_l2 X=1
ARGS="--workload C5 --pattern single --lanes 1"; run c5_l1_heavy_off X=1
ARGS="--workload C5 --pattern single --lanes 1"; run c5_l1_heavy_512 B2R_HEAVY=512
ARGS="--workload C5 --pattern single"; run c5_l4_heavy_off X=1
ARGS="--workload C5 --pattern single"; run c5_l4_heavy_512 B2R_HEAVY=512
ARGS="--workload C3 --pattern single"; run c3_l4 X=1
ARGS="--workload C2 --pattern single"; run c2_l4 X=1
