#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_group_count.py > gpurun_out/ab_group_count.txt 2>&1; cat gpurun_out/ab_group_count.txt
