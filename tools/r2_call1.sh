#!/bin/bash
# round-2 GPU call 1: parity suite (degenerate-scene property test on), bench under the driver's arguments
mkdir -p gpurun_out
FATESPLAT_PROPERTY_GPU=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench_driver.json 2> gpurun_out/c1_bench_driver.err
timeout 600 python bench.py --steps 200 --warmup 20 --quick > gpurun_out/c1_bench_long.json 2> gpurun_out/c1_bench_long.err
tail -3 gpurun_out/c1_pytest.log
cat gpurun_out/c1_bench_driver.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','kernels','speedup_vs_gpu_reference')})"
cat gpurun_out/c1_bench_long.json
