#!/bin/bash
mkdir -p gpurun_out
run() { # label lib env...
  label=$1; lib=$2; shift; shift
  out=$(env FATESPLAT_LIB=$lib "$@" timeout 300 python bench.py --steps 60 --warmup 10 --quick --no-extras --no-config3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print(round(d['ms_per_step'],4), {n:k[n]['us'] for n in ('blend_forward','blend_backward','tile_scan','tile_sort')})")
  echo "$label $out"
}
run base "" X=1
run g2 $PWD/fateavatar_b200/lib/variants/fwdg2.so X=1
run g8 $PWD/fateavatar_b200/lib/variants/fwdg8.so X=1
run g8_c2 $PWD/fateavatar_b200/lib/variants/fwdg8.so FATESPLAT_FWD_CTAS_PER_SM=2
