#!/bin/bash
for i in 1 2 3 4 5 6; do echo "== run $i"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | grep -E "passed|failed|AssertionError" | head -3; done
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mirror or config1" 2>&1 | grep -vE "^$" | tail -25
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mirror" 2>&1 | grep -vE "^$" | tail -25
