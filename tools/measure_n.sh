#!/bin/bash
# one multi-GPU bench line under the driver's arguments: bash tools/measure_n.sh N tag
mkdir -p gpurun_out
n=$1; tag=$2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-extras --no-config3 > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.err
echo "n=$n rc=$?"; tail -c 600 gpurun_out/${tag}_n$n.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_n$n.json').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','exchange_mode')}, 'e2e', d['e2e']['value'], 'check', (d.get('exchange_check') or {}).get('max_rel_err'), 'timing', d.get('exchange_timing'), d['clocks'])
except Exception as e: print('parse', e)
PY
