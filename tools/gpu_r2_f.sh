#!/bin/bash
# round 2, trip F
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_f.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_f.log
grep -E "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_f.log | cut -c1-250 | tail -30
timeout 300 python tools/time_proj.py > gpurun_out/time_proj3.txt 2>&1; cat gpurun_out/time_proj3.txt
timeout 300 python tools/time_attractor.py > gpurun_out/time_attractor3.txt 2>&1; grep -v diff gpurun_out/time_attractor3.txt
timeout 600 python tools/ab_switch.py USE_FUSED_PROJ_ANCHOR=1 USE_FUSED_PROJ_ANCHOR=0 > gpurun_out/ab_fused.txt 2>&1; cat gpurun_out/ab_fused.txt
timeout 300 python tools/lstm_profile.py 8 > gpurun_out/lstm_profile_b8_f.txt 2>&1; grep -A16 "backend 2" gpurun_out/lstm_profile_b8_f.txt
