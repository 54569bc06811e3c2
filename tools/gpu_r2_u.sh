#!/bin/bash
mkdir -p gpurun_out
DANET_LSTM_PROFILE=1 timeout 300 python tools/timeline.py > gpurun_out/timeline_phases_insitu.txt 2>&1; grep -A2 "period" gpurun_out/timeline_phases_insitu.txt | cut -c1-400
timeout 300 python tools/lstm_profile.py 32 2>&1 | grep -A14 "fp16 recurrent state" | head -20
