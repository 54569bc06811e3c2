#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_z.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_z.log
timeout 1200 python bench.py > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err; echo "bench exit $?"; tail -3 gpurun_out/bench_z.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_z.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'])
print('e2e', json.dumps(d['e2e'])[:700])
print('train', json.dumps(d['train'])[:900])
print('roofline', json.dumps(d['roofline'])[:300])
print('parity', json.dumps(d['parity'])[:300])
print('clocks', d['clocks'])
P
