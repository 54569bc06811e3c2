#!/bin/bash
# final N = 1 and N = 2 lines on the same box
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 --extras 0 --train-steps 10 > gpurun_out/bench_n2_final.json 2> gpurun_out/bench_n2_final.err
python - <<'PY'
import json
for f in ('bench_final3', 'bench_n2_final'):
    d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1]); t = d['train']
    print(f, 'value %.0f (%.3f ms)  e2e %.0f (lockstep %.0f)  train %.0f mixtures/s %.3f ms/step exposed AR %.3f ms launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['lockstep_value'], t['value'], t['ms_per_step'], t['allreduce_ms_exposed'], d['gpu_launches']))
PY
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 2>&1 | tail -1
