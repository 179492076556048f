#!/bin/bash
mkdir -p gpurun_out
bash tools/r2_pipe.sh "$@" 2>&1 | grep -v passed | tail -8
