#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lstm_seq_wide" > gpurun_out/pytest_t.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_t.log
timeout 300 python tools/lstm_wide_profile.py 32 > gpurun_out/wide_prof.txt 2>&1; cat gpurun_out/wide_prof.txt
