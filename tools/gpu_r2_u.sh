#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -s > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_final.log
grep -aE "^E  |passed|failed|exit|^FAILED" gpurun_out/pytest_final.log | cut -c1-250 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d['train']['value'], d['train']['ms_per_step'])
for o in d['other_configs']: print(o['config'][:50], round(o['ms_per_step_e2e'],3))
P
bash tools/cli_check.sh > gpurun_out/cli_check.log 2>&1; echo "cli_check exit $?"; tail -4 gpurun_out/cli_check.log | cut -c1-160
