#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2q_pytest.log
B="--steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-single --no-e2e"
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B $ARGS > gpurun_out/r2q_$name.log 2>&1; python - gpurun_out/r2q_$name.log $name <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; ok=True
        pk=r['per_kernel_ms']
        print(f"{sys.argv[2]:24s} value {d['value']:8.1f} fwd {pk.get('composite_fwd',0)*1e3:6.1f} bwd {pk.get('composite_bwd',0)*1e3:6.1f} us frac {r['frac']:.4f}")
        if r.get('per_view'):
            for k,v in r['per_view'].items(): print('   ', k, {w:(round(x['ms']*1e3,1), round(x['frac'],4)) for w,x in v.items()})
if not ok: print(sys.argv[2], 'FAILED'); print(open(sys.argv[1]).read()[-600:])
PY
}
ARGS="--workload C4"; run c4 X=1
ARGS="--workload C2 --pattern single"; run c2 X=1
ARGS="--workload C2 --pattern single --lanes 1 --frames 1"; run c2_solo X=1
ARGS="--workload C5 --pattern single"; run c5 X=1
ARGS="--workload C3 --pattern single"; run c3 X=1
