#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_optimise_loop_gpu.py -m gpu -x -q -l --tb=short > gpurun_out/c10_pytest.log 2>&1
echo "pytest rc=$?"; grep -n "^t  \|^sizes\|Error\|passed\|failed" gpurun_out/c10_pytest.log | head
