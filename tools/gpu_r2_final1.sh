#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s -k "fullsize" > gpurun_out/pytest_fullsize_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_fullsize_final.log
grep -E "max-norm errors|passed|failed|exit" gpurun_out/pytest_fullsize_final.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "not fullsize" > gpurun_out/pytest_rest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_rest_final.log; tail -2 gpurun_out/pytest_rest_final.log
timeout 900 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['gpu_launches_per_step'], d['train']['value'], d['train']['ms_per_step'])
PY
