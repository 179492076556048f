#!/bin/bash
mkdir -p gpurun_out
for arm in e2e value; do
FATESPLAT_BENCH_PROFILE=$arm ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_$arm.csv python bench.py --steps 4 --warmup 3 --no-extras --no-config3 > gpurun_out/r02_launches_$arm.log 2>&1
echo "$arm rc=$?"; wc -l gpurun_out/r02_launches_$arm.csv
done
