#!/bin/bash
# 2-GPU trip: NCCL sharded-gradient test, bench at N = 2 (inference + training with the bucketed exchange)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 500 > gpurun_out/pytest_multi2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_multi2.log; tail -5 gpurun_out/pytest_multi2.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --extras 0 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1800 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/copy_contention.py > gpurun_out/copy_n2.txt 2>&1; tail -8 gpurun_out/copy_n2.txt
