#!/bin/bash
# Usage: gpurun -- bash tools/gpu_quick.sh "<pytest -k expr>"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 120 -k "$1" > gpurun_out/pytest_quick.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_quick.log
grep -E "^E  |passed|failed|exit" gpurun_out/pytest_quick.log | cut -c1-300 | tail -40
