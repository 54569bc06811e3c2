#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_switch.py FLAGS_BY_MEMSET=1 FLAGS_BY_MEMSET=0 > gpurun_out/ab_memset.txt 2>&1; cat gpurun_out/ab_memset.txt
