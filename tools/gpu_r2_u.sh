#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "backward or gradients or lstm_tw" > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_u.log; grep -a "^E  " gpurun_out/pytest_u.log | head -5
ENC=lstm-orig timeout 600 python tools/time_train_groups.py > gpurun_out/time_train_lstm_orig.txt 2>&1; cat gpurun_out/time_train_lstm_orig.txt
