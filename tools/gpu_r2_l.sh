#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_switch.py STAGGER_US=0 STAGGER_US=25 STAGGER_US=40 STAGGER_US=55 > gpurun_out/ab_stagger.txt 2>&1; cat gpurun_out/ab_stagger.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:proj_anchor_kernel -s 3 -c 1 -o gpurun_out/prof_proj_anchor3 python tools/time_proj.py > gpurun_out/ncu_full_proj3.log 2>&1; tail -1 gpurun_out/ncu_full_proj3.log
