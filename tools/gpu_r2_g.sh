#!/bin/bash
# round 2, trip G: recurrent kernel after the peeled K16 slice (step period), prefetch depth 2 vs 3 in the whole step
mkdir -p gpurun_out
timeout 300 python tools/lstm_profile.py 8 > gpurun_out/lstm_profile_b8_g.txt 2>&1; grep -A16 "backend 2" gpurun_out/lstm_profile_b8_g.txt | head -20
timeout 300 python -m pytest tests -m gpu -q -k "lstm_seq or cfg2_emb" > gpurun_out/pytest_g.log 2>&1; tail -2 gpurun_out/pytest_g.log
timeout 600 python tools/ab_switch.py USE_FUSED_PROJ_ANCHOR=1 USE_FUSED_PROJ_ANCHOR=0 > gpurun_out/ab_depth2.txt 2>&1; cat gpurun_out/ab_depth2.txt
export DANET_NVCC_EXTRA=-DDANET_LSTM_PRE_DEPTH=3
timeout 600 python -c "import __graft_entry__ as g; g.build()" | tail -1
timeout 600 python tools/ab_switch.py USE_FUSED_PROJ_ANCHOR=1 USE_FUSED_PROJ_ANCHOR=0 > gpurun_out/ab_depth3.txt 2>&1; cat gpurun_out/ab_depth3.txt
export DANET_NVCC_EXTRA=-DDANET_LSTM_PRE_DEPTH=4
timeout 600 python -c "import __graft_entry__ as g; g.build()" | tail -1
timeout 600 python tools/ab_switch.py USE_FUSED_PROJ_ANCHOR=1 USE_FUSED_PROJ_ANCHOR=0 > gpurun_out/ab_depth4.txt 2>&1; cat gpurun_out/ab_depth4.txt
