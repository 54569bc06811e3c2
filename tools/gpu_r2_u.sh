#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_switch.py LSTM_LAUNCH_FIRST=0 LSTM_LAUNCH_FIRST=1 > gpurun_out/ab_launch_first.txt 2>&1; cat gpurun_out/ab_launch_first.txt
DANET_AB=LSTM_LAUNCH_FIRST=1 DANET_LSTM_PROFILE=2 timeout 300 python tools/timeline.py > gpurun_out/timeline_launch_first.txt 2>&1; grep "#\|total" gpurun_out/timeline_launch_first.txt | cut -c1-130
