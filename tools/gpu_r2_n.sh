#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_switch.py LSTM_HEADSTART_US=0 LSTM_HEADSTART_US=3 LSTM_HEADSTART_US=6 LSTM_HEADSTART_US=12 > gpurun_out/ab_headstart.txt 2>&1; cat gpurun_out/ab_headstart.txt
