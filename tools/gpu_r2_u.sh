#!/bin/bash
mkdir -p gpurun_out
DANET_AB=LSTM_HEADSTART_US=4 DANET_LSTM_PROFILE=1 timeout 300 python tools/timeline.py > gpurun_out/timeline_tm_hs4.txt 2>&1; grep "lstm\|gemm\|#" gpurun_out/timeline_tm_hs4.txt | head -24 | cut -c1-150
