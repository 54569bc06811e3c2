#!/bin/bash
# N = 2: bench (inference value / e2e / training in stream groups with one exchange after the sum) + the 2-rank NCCL tests
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --extras 0 --train-steps 10 > gpurun_out/bench_n2_v2.json 2> gpurun_out/bench_n2_v2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n2_v2.json').read().strip().splitlines()[-1]); t = d['train']
print('N=2 value %.0f (%.3f ms)  e2e %.0f (lockstep %.0f)  train %.0f mixtures/s %.3f ms/step exposed AR %.3f ms launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['lockstep_value'], t['value'], t['ms_per_step'], t['allreduce_ms_exposed'], d['gpu_launches']))
PY
tail -2 gpurun_out/bench_n2_v2.err | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 2>&1 | tail -2
