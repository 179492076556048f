#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/r2_pipe.sh "$@" 2>&1 | tail -6
timeout 300 python tools/determinism_check.py 40 2>&1 | tail -3
