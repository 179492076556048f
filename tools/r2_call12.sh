#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
echo "rc=$?"; tail -5 gpurun_out/c12_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c12_bench.json')); print({k:d.get(k) for k in ('value','ms_per_step','config3')}); print('e2e', d['e2e']['value'], 'gpu_ref', d['gpu_reference'])
PY
