#!/bin/bash
# 2-GPU trip b: NCCL test again, NCCL channel count vs training step, training precision tool
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 > gpurun_out/pytest_multi2b.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi2b.log; tail -3 gpurun_out/pytest_multi2b.log | cut -c1-300
for CH in 0 2 4 8; do
  if [ "$CH" != "0" ]; then export NCCL_MAX_NCHANNELS=$CH; export NCCL_MIN_NCHANNELS=1; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$CH bench.py --gpus 2 --steps 3 --warmup 3 --extras 0 --train-steps 10 > gpurun_out/bench_n2_ch$CH.json 2> gpurun_out/bench_n2_ch$CH.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_n2_ch$CH.json').read().strip().splitlines()[-1])
t = d['train']
print('NCCL_MAX_NCHANNELS=$CH train %.0f mixtures/s  %.3f ms/step  exposed all-reduce %.3f ms  bptt %.3f ms' % (t['value'], t['ms_per_step'], t['allreduce_ms_exposed'], t['roofline']['ms_per_launch']))
PY
done
unset NCCL_MAX_NCHANNELS NCCL_MIN_NCHANNELS
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/train_precision.py > gpurun_out/train_precision.txt 2>&1; cat gpurun_out/train_precision.txt
