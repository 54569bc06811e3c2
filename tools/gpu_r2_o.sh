#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -k "pipelined or fused_inference or separate" > gpurun_out/pytest_o.log 2>&1; tail -2 gpurun_out/pytest_o.log
timeout 900 python tools/ab_switch.py FLAGS_BY_MEMSET=1 FLAGS_BY_MEMSET=0 > gpurun_out/ab_zero_kernel.txt 2>&1; cat gpurun_out/ab_zero_kernel.txt
