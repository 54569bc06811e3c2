#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:proj_anchor_kernel -s 3 -c 1 -o gpurun_out/prof_proj_anchor2 python tools/time_proj.py > gpurun_out/ncu_full_proj2.log 2>&1; tail -2 gpurun_out/ncu_full_proj2.log
ls -la gpurun_out/prof_proj_anchor2.ncu-rep
