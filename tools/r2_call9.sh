#!/bin/bash
mkdir -p gpurun_out
FATESPLAT_PROPERTY_GPU=1 timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/c9_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "frame golden|passed|failed|rror" gpurun_out/c9_pytest.log | tail -20
