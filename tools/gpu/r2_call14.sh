#!/bin/bash
# register-budget variants of the composites (tuning builds in gpurun_variants/, loaded with B2R_LIB)
cd "$GRAFT_REPO_ROOT"
B="--steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-single --no-e2e"
run() { name=$1; shift; env "$@" timeout 600 python bench.py $B $ARGS > gpurun_out/r2n_$name.log 2>&1; python - gpurun_out/r2n_$name.log $name <<'PY'
import json,sys
ok=False
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; ok=True
        pk=r['per_kernel_ms']
        print(f"{sys.argv[2]:24s} value {d['value']:8.1f} fwd {pk.get('composite_fwd',0)*1e3:6.1f} bwd {pk.get('composite_bwd',0)*1e3:6.1f} us")
if not ok: print(sys.argv[2], 'FAILED'); print(open(sys.argv[1]).read()[-600:])
PY
}
for v in default fb64 fb64_r64; do
  if [ $v = default ]; then L=exavatar_release_b200/libb200raster.so; else L=gpurun_variants/lib_$v.so; fi
  ARGS="--workload C4"; run c4_$v B2R_LIB=$PWD/$L
  ARGS="--workload C2 --pattern single"; run c2_$v B2R_LIB=$PWD/$L
done

