#!/bin/bash
# round 2, trip A: the whole GPU suite (incl. the new full-size gates), recurrent-kernel fixed cost + phase stamps, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -s -k "fullsize" > gpurun_out/pytest_fullsize.log 2>&1
echo "pytest fullsize exit $?" >> gpurun_out/pytest_fullsize.log
grep -E "max-norm errors|^E  |passed|failed|exit" gpurun_out/pytest_fullsize.log | cut -c1-400 | tail -40
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "not fullsize" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|exit" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -20
timeout 300 python tools/lstm_fixed_cost.py > gpurun_out/lstm_fixed_cost.txt 2>&1; cat gpurun_out/lstm_fixed_cost.txt
timeout 300 python tools/lstm_fixed_cost.py 8 > gpurun_out/lstm_fixed_cost_b8.txt 2>&1; cat gpurun_out/lstm_fixed_cost_b8.txt
timeout 300 python tools/lstm_profile.py 8 > gpurun_out/lstm_profile_b8.txt 2>&1; tail -40 gpurun_out/lstm_profile_b8.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -c 3000 gpurun_out/bench_a.json; tail -5 gpurun_out/bench_a.err
