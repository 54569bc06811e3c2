#!/bin/bash
# round 2, trip C: new anchor kernel (tests + timing), H2D prefetch (timeline), bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_c.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_c.log
grep -E "^E  |passed|failed|exit" gpurun_out/pytest_c.log | cut -c1-300 | tail -20
timeout 300 python tools/time_attractor.py > gpurun_out/time_attractor.txt 2>&1; cat gpurun_out/time_attractor.txt
timeout 300 python tools/timeline.py 32000 host > gpurun_out/timeline_host2.txt 2>&1; grep -E "start|end|total" gpurun_out/timeline_host2.txt
timeout 300 python tools/timeline.py 32000 > gpurun_out/timeline_dev2.txt 2>&1; tail -16 gpurun_out/timeline_dev2.txt
timeout 900 python bench.py --steps 10 --train-steps 5 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e'], d['parity'], d['precision_ab'])
print([(k['kernel'], round(k.get('ms_per_launch') or k.get('ms_per_step'), 4), round(k['frac'], 3)) for k in d['kernels']])
print(d['train']['value'], d['train']['ms_per_step'])
PY
tail -5 gpurun_out/bench_c.err
