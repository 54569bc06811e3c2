#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "proj_anchor or fused_inference or fullsize or full_size" > gpurun_out/pytest_k.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_k.log
grep -E "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_k.log | cut -c1-250 | tail -12
timeout 300 python tools/time_proj.py > gpurun_out/time_proj4.txt 2>&1; cat gpurun_out/time_proj4.txt
timeout 600 python tools/ab_switch.py USE_FUSED_PROJ_ANCHOR=1 USE_FUSED_PROJ_ANCHOR=0 > gpurun_out/ab_fused2.txt 2>&1; cat gpurun_out/ab_fused2.txt
