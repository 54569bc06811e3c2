#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/time_gemm_tn.txt
for tn in 128 256; do DANET_GEMM_TN=$tn timeout 300 python tools/time_gemm_tn.py >> gpurun_out/time_gemm_tn.txt 2>&1; done
cat gpurun_out/time_gemm_tn.txt
