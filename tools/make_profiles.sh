#!/bin/bash
# Runs on the GPU box (under gpurun): ncu launch list of the bench command + full captures of the main kernels.
set -x
R=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --quick --steps 4 --warmup 3 > gpurun_out/${R}_launches.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"blend_backward|blend_forward_kernel|tile_sort_kernel|preprocess_kernel|preprocess_backward_kernel|scatter_kernel|tile_scan_kernel|pose_forward_kernel|pose_backward_kernel|flame_" \
    -s 61 -c 15 -o gpurun_out/${R}_kernels python bench.py --quick --steps 4 --warmup 3 > gpurun_out/${R}_full.log 2>&1
ls -la gpurun_out
