#!/bin/bash
# quick knob sweep: per-kernel times from bench.py --quick under different tuning env vars
mkdir -p gpurun_out
for cfg in "FATESPLAT_BWD_CTAS_PER_SM=2" "FATESPLAT_BWD_CTAS_PER_SM=3" "FATESPLAT_BWD_CTAS_PER_SM=4" "FATESPLAT_BWD_CTAS_PER_SM=5"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 100 --warmup 10 --quick 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print(round(d['ms_per_step'],4), {n:k[n]['us'] for n in ('blend_forward','blend_backward','tile_scan','scatter','tile_sort')})"
done 2>&1 | tee gpurun_out/c2_knobs.log
