#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "pipelined or fused_inference or separate or full_size or cfg2 or streaming" > gpurun_out/pytest_p.log 2>&1; tail -2 gpurun_out/pytest_p.log
timeout 900 python tools/ab_switch.py FLAG_SETS=8 FLAG_SETS=0 > gpurun_out/ab_flagpool.txt 2>&1; cat gpurun_out/ab_flagpool.txt
