#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "lstm_seq or pipelined or programmatic or cfg2_emb or golden" > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_u.log
timeout 300 python tools/lstm_profile.py 32 2>&1 | grep -A14 "fp16 recurrent state" | head -18
timeout 900 python tools/ab_switch.py PROGRAMMATIC_LSTM_LAUNCH=1 > gpurun_out/ab_arm.txt 2>&1; cat gpurun_out/ab_arm.txt
