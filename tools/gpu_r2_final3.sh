#!/bin/bash
# Final evidence trip of round 2 on ONE B200 (final code): tests, smoke, bench (both arms), ncu launch list, ncu --set full of
# the recurrent kernels, timelines, encoder rates.  Summarised into profiles/r02_* afterwards.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_final.log
grep -aE "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_final.log | cut -c1-250 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 700 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
tail -c 400 gpurun_out/bench_ref_final.json
timeout 300 python tools/timeline.py 32000 > gpurun_out/timeline_final_dev.txt 2>&1; tail -1 gpurun_out/timeline_final_dev.txt
timeout 300 python tools/timeline.py 32000 host > gpurun_out/timeline_final_host.txt 2>&1; tail -1 gpurun_out/timeline_final_host.txt
timeout 300 python tools/lstm_profile.py 8 > gpurun_out/lstm_profile_final.txt 2>&1; grep "==" gpurun_out/lstm_profile_final.txt
timeout 600 python tools/encoder_rates.py > gpurun_out/encoder_rates_final.txt 2>&1; tail -4 gpurun_out/encoder_rates_final.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graph 0 --train-steps 0 --extras 0 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph 0 --train-steps 0 --extras 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_tc2 -s 8 -c 2 -o gpurun_out/prof_final_lstm $B > gpurun_out/ncu_f1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_wide_kernel -s 2 -c 1 -o gpurun_out/prof_final_lstm_wide python tools/lstm_wide_profile.py 8 > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out/prof_final_lstm*.ncu-rep
