#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "wide or backward and 600 or backward and 416 or lstm_tw or lstm_orig" > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_u.log
timeout 300 python tools/lstm_wide_bwd_profile.py 32 > gpurun_out/wide_bwd_prof.txt 2>&1; cat gpurun_out/wide_bwd_prof.txt
timeout 300 python tools/lstm_wide_profile.py 32 2>&1 | head -7
