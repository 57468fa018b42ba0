#!/bin/bash
# A/B of library variants (gpurun_variants/<name>.so, loaded through B2R_LIB): per-view kernel times of one C4 training
# frame (serial), the C4 five-render value leg and the C2 single-render value leg; then the full GPU suite on the product
# build.  usage: ab_variants.sh <tag> <variant>...
cd "$GRAFT_REPO_ROOT"
tag=$1; shift
summ() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('   value %.1f  ms/step %.3f  per_kernel %s' % (d['value'], d['ms_per_step'], json.dumps(r['per_kernel_ms'])))
PY
}
for v in "$@"; do
  echo "== $v"
  B2R_LIB=$PWD/gpurun_variants/$v.so timeout 300 python tools/five_breakdown.py > gpurun_out/${tag}_five_$v.txt 2>&1
  grep -E "view|project \+ bin" gpurun_out/${tag}_five_$v.txt | head -7
  B2R_LIB=$PWD/gpurun_variants/$v.so timeout 400 python bench.py --steps 20 --warmup 5 --no-e2e --no-eager --no-single --no-cpu-baseline > gpurun_out/${tag}_c4_$v.log 2>&1; summ gpurun_out/${tag}_c4_$v.log
  B2R_LIB=$PWD/gpurun_variants/$v.so timeout 400 python bench.py --workload C2 --pattern single --steps 20 --warmup 5 --no-e2e --no-eager --no-cpu-baseline > gpurun_out/${tag}_c2_$v.log 2>&1; summ gpurun_out/${tag}_c2_$v.log
done
rm -f gpurun_out/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${tag}_pytest.log
cp gpurun_out/parity_report.jsonl gpurun_out/${tag}_parity.jsonl 2>/dev/null
