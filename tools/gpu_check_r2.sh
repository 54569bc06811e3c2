#!/bin/bash
# Round-2 evidence trip on ONE B200: tests, smoke, bench (both arms), ncu launch list, ncu --set full captures, timelines,
# BASELINE.md table rows, compute-sanitizer.  Everything lands in gpurun_out/ and is summarised into profiles/r02_* afterwards.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt; lscpu | grep "Model name" >> gpurun_out/smi.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_final.log
grep -E "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_final.log | cut -c1-250 | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 1500 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
tail -c 600 gpurun_out/bench_ref_final.json
timeout 600 python tools/baseline_table.py > gpurun_out/baseline_table.json 2> gpurun_out/baseline_table.err; cat gpurun_out/baseline_table.json; tail -2 gpurun_out/baseline_table.err
timeout 300 python tools/timeline.py 32000 > gpurun_out/timeline_final_dev.txt 2>&1; tail -1 gpurun_out/timeline_final_dev.txt
timeout 300 python tools/timeline.py 32000 host > gpurun_out/timeline_final_host.txt 2>&1; tail -1 gpurun_out/timeline_final_host.txt
timeout 300 python tools/lstm_profile.py 8 > gpurun_out/lstm_profile_final.txt 2>&1; grep "==" gpurun_out/lstm_profile_final.txt
timeout 300 python tools/time_proj.py > gpurun_out/time_proj_final.txt 2>&1
timeout 300 python tools/time_attractor.py > gpurun_out/time_attractor_final.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --train-steps 0 --extras 0 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph 0 --train-steps 0 --extras 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc2 -s 8 -c 2 -o gpurun_out/prof_final_lstm $B > gpurun_out/ncu_f1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:proj_anchor_kernel -s 3 -c 1 -o gpurun_out/prof_final_proj $B > gpurun_out/ncu_f2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 12 -c 2 -o gpurun_out/prof_final_gemm $B > gpurun_out/ncu_f3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_kernel -s 2 -c 1 -o gpurun_out/prof_final_stft $B > gpurun_out/ncu_f4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mask_istft_kernel -s 2 -c 1 -o gpurun_out/prof_final_k4 $B > gpurun_out/ncu_f5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:anchor2_mma -s 2 -c 1 -o gpurun_out/prof_final_attr python tools/time_attractor.py > gpurun_out/ncu_f6.log 2>&1
ls gpurun_out/prof_final_*.ncu-rep
bash tools/sanitize.sh 2>&1 | grep -E "exit|passed|failed|SUMMARY"
