#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_v.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_v.log
timeout 600 python tools/time_train.py > gpurun_out/time_train_v.txt 2>&1; tail -12 gpurun_out/time_train_v.txt
