#!/bin/bash
# per-CTA timelines of the composites (trace build gpurun_variants/trace.so)
cd "$GRAFT_REPO_ROOT"
for wl in C2 C4; do
  B2R_LIB=$PWD/gpurun_variants/trace.so timeout 300 python tools/cta_trace.py --workload $wl > gpurun_out/$1_cta_$wl.txt 2>&1; echo "$wl rc=$?"
done
