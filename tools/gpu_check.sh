#!/bin/bash
# One GPU-box trip: parity tests, smoke, bench (both backends), ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 $PYTEST_FLAGS > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^E  |passed|failed|exit" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -3 gpurun_out/smoke.log
for BE in 1; do
timeout 600 python bench.py --backend $BE --steps 10 --warmup 3 > gpurun_out/bench_b$BE.json 2> gpurun_out/bench_b$BE.err
tail -c 2500 gpurun_out/bench_b$BE.json; tail -5 gpurun_out/bench_b$BE.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_b1.csv python bench.py --backend 1 --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --train-steps 0 --extras 0 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
# one full-metric capture of the dominant kernel (3 launches), brought back as a report
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lstm_tc -s 8 -c 2 -o gpurun_out/prof_lstm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph 0 --train-steps 0 --extras 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 20 -c 2 -o gpurun_out/prof_gemm python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph 0 --train-steps 0 --extras 0 > gpurun_out/ncu_full2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attractor_partial -s 2 -c 1 -o gpurun_out/prof_attr python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph 0 --train-steps 0 --extras 0 > gpurun_out/ncu_full3.log 2>&1
for KN in stft_kernel mask_istft_kernel; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KN -s 2 -c 1 -o gpurun_out/prof_$KN python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph 0 --train-steps 0 --extras 0 > gpurun_out/ncu_$KN.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
