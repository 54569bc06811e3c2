#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/overlap_steps.py 2 > gpurun_out/overlap_steps.txt 2>&1; cat gpurun_out/overlap_steps.txt
timeout 600 python tools/overlap_steps.py 3 >> gpurun_out/overlap_steps.txt 2>&1; tail -3 gpurun_out/overlap_steps.txt
