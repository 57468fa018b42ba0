#!/bin/bash
# round 2, call 2: segmented composites (v4) + merged five-render plan: parity, A/B timing
cd "$GRAFT_REPO_ROOT"
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest_v4.log 2>&1; echo "pytest v4 rc=$?" | tee -a gpurun_out/r2b_pytest_v4.log
cp gpurun_out/parity_report.jsonl gpurun_out/r2b_parity_v4.jsonl 2>/dev/null
B2R_COMPOSITE=v3 timeout 1200 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s -k "not merged" > gpurun_out/r2b_pytest_v3.log 2>&1; echo "pytest v3 rc=$?" | tee -a gpurun_out/r2b_pytest_v3.log
for eng in merged separate; do
  timeout 600 python bench.py --steps 10 --warmup 3 --engine $eng --no-cpu-baseline --no-eager --no-single > gpurun_out/r2b_bench_c4_$eng.log 2>&1; echo "bench $eng rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 3 --workload C2 --pattern single --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_c2_v4.log 2>&1
B2R_COMPOSITE=v3 timeout 600 python bench.py --steps 20 --warmup 3 --workload C2 --pattern single --no-cpu-baseline --no-e2e > gpurun_out/r2b_bench_c2_v3.log 2>&1
grep -h "passed\|failed" gpurun_out/r2b_pytest_v4.log gpurun_out/r2b_pytest_v3.log | tail -4
for f in gpurun_out/r2b_bench_*.log; do echo $f; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print(' value',round(d['value'],1),'e2e',d['e2e'] and round(d['e2e']['value'],1),'frac',round(r['frac'],4),r['kernel'],r['per_kernel_ms'])
    elif 'rror' in l: print(l[:300])
PY
done
