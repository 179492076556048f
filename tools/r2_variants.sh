#!/bin/bash
mkdir -p gpurun_out
run() { # label lib env...
  label=$1; lib=$2; shift; shift
  out=$(env FATESPLAT_LIB=$lib "$@" timeout 300 python bench.py --steps 60 --warmup 10 --quick --no-extras --no-config3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print(round(d['ms_per_step'],4), {n:k[n]['us'] for n in k})")
  echo "$label $out"
}
run base_4096_512 "" X=1
run sort2048 $PWD/fateavatar_b200/lib/variants/sort2048.so X=1
run sort4096_256 $PWD/fateavatar_b200/lib/variants/sort4096_256.so X=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/gpu_check.py c5 2>&1 | grep -E "stage us|new\(async\)|ref  fwd"
