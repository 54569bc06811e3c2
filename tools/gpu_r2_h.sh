#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_h.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_h.log
grep -E "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_h.log | cut -c1-250 | tail -30
timeout 300 python tools/lstm_profile.py 8 > gpurun_out/lstm_profile_b8_h.txt 2>&1; grep "backend 2" gpurun_out/lstm_profile_b8_h.txt
