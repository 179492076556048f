#!/bin/bash
# A/B of blend-backward variants: per-kernel times at config 2 and config 5, then parity of the last variant
mkdir -p gpurun_out
V=$PWD/fateavatar_b200/lib/variants
run() { # label lib env...
  label=$1; lib=$2; shift; shift
  out=$(env FATESPLAT_LIB=$lib "$@" timeout 300 python bench.py --steps 60 --warmup 10 --quick --no-extras --no-config3 2>gpurun_out/pipe_$label.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print(round(d['ms_per_step'],4), {n:k[n]['us'] for n in ('blend_forward','blend_backward','tile_scan')})")
  c5=$(env FATESPLAT_LIB=$lib "$@" timeout 300 python tools/gpu_check.py c5 2>&1 | grep -E "stage us" | python -c "
import sys,ast
d=ast.literal_eval(sys.stdin.read().split('stage us:')[1].strip()); print({k:d[k] for k in ('blend_forward','blend_backward')})")
  echo "$label $out c5 $c5"
}
run base "" X=1
for v in "$@"; do run $v $V/$v.so X=1; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
