#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "mask or istft or separate or fused or cfg4 or cfg2_emb or cfg5 or golden" > gpurun_out/pytest_u.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_u.log
timeout 300 python tools/time_kernels.py > gpurun_out/time_kernels_u.txt 2>&1; grep -i "stft\|mask" gpurun_out/time_kernels_u.txt | head
