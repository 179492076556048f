#!/bin/bash
# round-2 GPU call 3 (2 GPUs): sharded step at N=1 and N=2
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err
echo "n1 rc=$?"; tail -3 gpurun_out/c3_bench_n1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/c3_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','step_issue','gpu_launches_per_step')})
except Exception as e: print('n1 parse', e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c3_bench_n2.json 2> gpurun_out/c3_bench_n2.err
echo "n2 rc=$?"; tail -5 gpurun_out/c3_bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/c3_bench_n2.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','exchange_check','exchange_mode')})
except Exception as e: print('n2 parse', e)
PY
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
