#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "backward or gradients or lstm_seq" > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_u.log
