#!/bin/bash
mkdir -p gpurun_out
bash tools/cli_check.sh > gpurun_out/cli_check.log 2>&1; echo "cli_check exit $?"; tail -12 gpurun_out/cli_check.log | cut -c1-200
cat > /tmp/conv.json <<'J'
{"ENCODER_TYPE": "conv-bilstm-v1", "TRAIN_ESTIMATOR_METHOD": "anchor", "INFER_ESTIMATOR_METHOD": "anchor", "SEPARATOR_TYPE": "dot-softmax-orig", "LR_DECAY_TYPE": "fixed", "NUM_EPOCH_PER_LR_DECAY": 1}
J
timeout 300 python main.py -m train -ne 2 -bs 2 -c /tmp/conv.json --no-save-on-epoch > gpurun_out/cli_conv_train.log 2>&1; echo "conv train exit $?"; tail -6 gpurun_out/cli_conv_train.log | cut -c1-200
