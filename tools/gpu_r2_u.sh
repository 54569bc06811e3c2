#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "pipelined or programmatic or fused_inference or separate or full_size or cfg2 or streaming or golden" > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_u.log
timeout 900 python tools/ab_switch.py PROGRAMMATIC_LSTM_LAUNCH=1 PROGRAMMATIC_LSTM_LAUNCH=0 > gpurun_out/ab_pdl.txt 2>&1; cat gpurun_out/ab_pdl.txt
DANET_LSTM_PROFILE=2 timeout 300 python tools/timeline.py > gpurun_out/timeline_pdl.txt 2>&1; grep "#\|total" gpurun_out/timeline_pdl.txt | cut -c1-130
