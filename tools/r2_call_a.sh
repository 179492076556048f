#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-config3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -c 300 gpurun_out/a_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/a_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),'clocks',d['clocks'])
print({k:v['us'] for k,v in d['kernels'].items()})
P
timeout 300 python tools/gpu_check.py c2 c5 2>&1 | grep -E "stage us|new\(async\)|ref  fwd|ratio" 
