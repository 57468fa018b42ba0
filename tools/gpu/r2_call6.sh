#!/bin/bash
# round 2, call 6: PDL chain + self-cleaning ctx + striding backward grid: parity, A/B, ncu
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2e_pytest.log
grep -h "passed\|failed\|^FAILED" gpurun_out/r2e_pytest.log | tail -5
B="--steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-single --no-e2e"
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B $ARGS > gpurun_out/r2e_$name.log 2>&1; python - gpurun_out/r2e_$name.log $name <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; ok=True
        pk=r['per_kernel_ms']
        print(f"{sys.argv[2]:24s} value {d['value']:8.1f} ms/step {d['ms_per_step']:.4f} fwd {pk.get('composite_fwd',0)*1e3:6.1f} bwd {pk.get('composite_bwd',0)*1e3:6.1f} us frac {r['frac']:.4f} ({r['kernel'][10:]}) sum {r.get('sum_kernel_ms_per_training_frame', r.get('sum_kernel_ms_per_frame',0))*1e3:7.1f}us", {k:round(v*1e3,1) for k,v in pk.items() if not k.startswith('composite')})
if not ok: print(sys.argv[2], 'FAILED'); print(open(sys.argv[1]).read()[-800:])
PY
}
ARGS="--workload C2 --pattern single --lanes 1 --frames 1"; run c2_solo_pdl X=1
ARGS="--workload C2 --pattern single --lanes 1 --frames 1"; run c2_solo_nopdl B2R_PDL=0
ARGS="--workload C2 --pattern single"; run c2_l4_pdl X=1
ARGS="--workload C2 --pattern single"; run c2_l4_nopdl B2R_PDL=0
ARGS="--workload C4 --engine merged"; run c4_merged_l3_pdl X=1
ARGS="--workload C4 --engine merged"; run c4_merged_l3_nopdl B2R_PDL=0
ncu --set full --clock-control none --import-source on -k regex:composite -s 4 -c 2 -o gpurun_out/prof_r02b python tools/profile_frame.py --workload C2 --frames 3 > gpurun_out/ncu_r02b.log 2>&1
ls -la gpurun_out/prof_r02b.ncu-rep
