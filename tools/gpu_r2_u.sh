#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s -k "proj_anchor or fused_inference or cfg2_emb or three_layer or golden or separate" > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; grep -a "max-norm\|passed\|failed" gpurun_out/pytest_u.log | tail -8
timeout 300 python tools/time_proj.py > gpurun_out/time_proj_u.txt 2>&1; cat gpurun_out/time_proj_u.txt
timeout 900 python tools/ab_switch.py PROGRAMMATIC_LSTM_LAUNCH=1 > gpurun_out/ab_proj.txt 2>&1; cat gpurun_out/ab_proj.txt
