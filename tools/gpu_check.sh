#!/bin/bash
# One GPU-box trip: parity tests, smoke, bench.  Usage: gpurun -- bash tools/gpu_check.sh [backend]
BACKEND=${1:-0}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 $PYTEST_FLAGS > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --backend $BACKEND --steps 10 --warmup 3 > gpurun_out/bench_b$BACKEND.json 2> gpurun_out/bench_b$BACKEND.err
tail -c 3000 gpurun_out/bench_b$BACKEND.json; tail -5 gpurun_out/bench_b$BACKEND.err
