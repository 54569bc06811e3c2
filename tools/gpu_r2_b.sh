#!/bin/bash
# round 2, trip B: changed tests, bench with the restructured line, reference arm smoke, timelines (device / e2e / half length)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s -k "fullsize or multi or packed" > gpurun_out/pytest_b.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_b.log
grep -E "max-norm errors|^E  |passed|failed|exit" gpurun_out/pytest_b.log | cut -c1-300 | tail -30
timeout 900 python bench.py > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
tail -c 2500 gpurun_out/bench_b.json; tail -5 gpurun_out/bench_b.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 1500 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
timeout 300 python tools/timeline.py 32000 > gpurun_out/timeline_dev.txt 2>&1; tail -18 gpurun_out/timeline_dev.txt
timeout 300 python tools/timeline.py 32000 host > gpurun_out/timeline_host.txt 2>&1; tail -18 gpurun_out/timeline_host.txt
timeout 300 python tools/timeline.py 16000 > gpurun_out/timeline_half.txt 2>&1; tail -18 gpurun_out/timeline_half.txt
