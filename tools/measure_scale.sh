#!/bin/bash
# scaling run under the driver's arguments: N = 1, 2, 4, 8 (as many as the box has)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -gt $NG ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${1}_n$n.json 2> gpurun_out/${1}_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${1}_n$n.json 2> gpurun_out/${1}_n$n.err
  fi
  echo "n=$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${1}_n$n.json')); print({k:d.get(k) for k in ('value','ms_per_step','exchange_mode')}, 'e2e', d['e2e']['value'], 'exch_us', d['kernels'].get('exchange'), 'check', (d.get('exchange_check') or {}).get('max_rel_err'), 'timing', d.get('exchange_timing'))
except Exception as e: print('parse', e)
PY
done
