#!/bin/bash
# N=1: parity suite, ncu launch list + full captures (profiles), full bench line under the driver's arguments
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/make_profiles.sh r02 > gpurun_out/r02_make_profiles.log 2>&1
bash tools/profile_arms.sh > gpurun_out/prof_e2e.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/gpu_check.py c2 c5 > gpurun_out/r02_gpu_check.log 2>&1; grep -E "stage us|new\(async\)|ref  fwd" gpurun_out/r02_gpu_check.log
cp gpurun_out/gpu_check.json gpurun_out/r02_gpu_check_c2_c5.json
