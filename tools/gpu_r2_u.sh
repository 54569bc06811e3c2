#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "wide or lstm_orig" > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_u.log
timeout 300 python tools/lstm_wide_profile.py 32 > gpurun_out/wide_prof.txt 2>&1; cat gpurun_out/wide_prof.txt
timeout 900 python bench.py --train-steps 0 --no-cpu-baseline > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_u.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'])
for o in d['other_configs']: print(o['config'][:60], o['ms_per_step_e2e'], o['mixtures_per_s_e2e'])
P
