#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s -k "lstm_seq_wide or lstm_orig" > gpurun_out/pytest_t.log 2>&1; echo "pytest exit $?"; grep -a "max-norm\|passed\|failed\|Error" gpurun_out/pytest_t.log | tail
timeout 300 python tools/lstm_wide_profile.py 32 > gpurun_out/wide_prof.txt 2>&1; cat gpurun_out/wide_prof.txt
timeout 600 python tools/encoder_rates.py > gpurun_out/encoder_rates_s.txt 2>&1; tail -4 gpurun_out/encoder_rates_s.txt
