#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
echo "rc=$?"; grep -E "Elapsed|Error|error" gpurun_out/c12_bench.err | tail -5
python - <<'PY'
import json
d=json.load(open('gpurun_out/c12_bench.json')); print({k:d.get(k) for k in ('value','ms_per_step','config3')}); print('e2e', d['e2e']['value'])
x=d['extras']
if 'error' in x: print(x)
else:
    print('c1', x['config1']); c5=dict(x['config5']); pv=c5.pop('per_view'); print('c5', c5); print(pv[0], pv[9]); print('knn', x['knn'])
PY
