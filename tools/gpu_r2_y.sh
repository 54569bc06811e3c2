#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "pit or stream_groups or gradients or sharded or train or clip or golden" > gpurun_out/pytest_y.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_y.log
timeout 600 python tools/time_train.py > gpurun_out/time_train_y.txt 2>&1; tail -14 gpurun_out/time_train_y.txt
timeout 600 python tools/time_train_groups.py > gpurun_out/time_train_groups.txt 2>&1; cat gpurun_out/time_train_groups.txt
