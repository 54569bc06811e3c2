#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/lstm_wide_bwd_profile.py 32 > gpurun_out/wide_bwd_prof.txt 2>&1; cat gpurun_out/wide_bwd_prof.txt
