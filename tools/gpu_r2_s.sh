#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lstm_seq" > gpurun_out/pytest_s.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_s.log
timeout 600 python tools/encoder_rates.py > gpurun_out/encoder_rates_s.txt 2>&1; tail -8 gpurun_out/encoder_rates_s.txt
