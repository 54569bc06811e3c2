#!/bin/bash
# 8-GPU trip: bench at N = 8 and N = 4 (inference value / e2e / training in stream groups with one exchange after the sum)
mkdir -p gpurun_out
for N in 8 4; do
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 --extras 0 --train-steps 10 > gpurun_out/bench_n${N}_v2.json 2> gpurun_out/bench_n${N}_v2.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n${N}_v2.json').read().strip().splitlines()[-1]); t = d['train']
print('N=$N value %.0f (%.3f ms)  e2e %.0f (lockstep %.0f)  train %.0f mixtures/s %.3f ms/step exposed AR %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['lockstep_value'], t['value'], t['ms_per_step'], t['allreduce_ms_exposed']))
PY
tail -2 gpurun_out/bench_n${N}_v2.err | cut -c1-200
done
