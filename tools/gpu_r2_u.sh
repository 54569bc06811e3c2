#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d['train']['value'], d['train']['ms_per_step'])
print([ (k['kernel'], round(k['frac'],3)) for k in d['kernels']], d['roofline']['frac'], d['clocks'])
P
timeout 600 python tools/encoder_rates.py > gpurun_out/encoder_rates_final.txt 2>&1; tail -4 gpurun_out/encoder_rates_final.txt
