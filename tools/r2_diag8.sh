#!/bin/bash
# N-GPU diagnosis of the sharded step: variants of the exchange, device-resident arm only
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
run() {  # name, env..., extra args after --
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 40 --warmup 5 --quick "$@" > gpurun_out/d8_$name.json 2> gpurun_out/d8_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/d8_$name.json')); print('$name', round(d['ms_per_step'],4), round(d['value']), d.get('exchange_timing'))
except Exception as e: print('$name parse', e)
PY
}




run trace FATESPLAT_EXCHANGE_TRACE=1 --
