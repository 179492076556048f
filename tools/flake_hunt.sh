#!/bin/bash
# flake hunt: the parity file N times in fresh processes; prints every failure with its diagnostics
mkdir -p gpurun_out
fail=0
for i in $(seq 1 ${1:-20}); do
  out=$(timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_graph_gpu.py -m gpu -x -q 2>&1)
  if ! echo "$out" | grep -q " passed"; then echo "== run $i: no pass line"; echo "$out" | tail -30; fail=$((fail+1)); continue; fi
  if echo "$out" | grep -q "failed"; then echo "== run $i FAILED"; echo "$out" | grep -E "AssertionError|worst_pixel|FAILED|Error" | head -12; fail=$((fail+1)); fi
done
echo "runs ${1:-20} failures $fail"
