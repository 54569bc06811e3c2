#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 120 -x -k "pipelined_input" > gpurun_out/pytest_i0.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_i0.log
grep -E "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_i0.log | cut -c1-250 | tail -12
timeout 600 python tools/ab_switch.py PIPELINE_INPUT_GEMM=1 PIPELINE_INPUT_GEMM=0 > gpurun_out/ab_pipe.txt 2>&1; cat gpurun_out/ab_pipe.txt
timeout 300 python tools/timeline.py 32000 > gpurun_out/timeline_dev3.txt 2>&1; tail -16 gpurun_out/timeline_dev3.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_i.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_i.log
grep -E "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_i.log | cut -c1-250 | tail -12
