#!/bin/bash
# compute-sanitizer over a small slice of the GPU tests (memcheck + racecheck + synccheck)
mkdir -p gpurun_out
SEL='test_stft_vs_oracle and 777 or test_istft_vs_oracle and 33 or test_center or test_mix_features or test_linear and 37 or test_lstm_seq and 2-3-12 or test_attractor_anchor and 3-2-9 or test_attractor_truth and 3-2-9 or test_mask_cmul and 3-2-9 or test_pit_mse and 4-2-7 or test_head_backward and 3-2-9 or test_lstm_layer_backward and 3-9-40 or test_clip_adam or test_lstm_seq_packed_weights and 2-8-40 or test_mask_cmul_istft_fused and 2-2-40 or test_conv2d_maxpool_add and 16-32 or test_split_operand_paired or test_model_forward_golden and convbilstm and 1 or test_model_gradients_golden and c3 or test_proj_anchor_fused and 2-5-64 or test_proj_anchor_fused and 2-128-100 or test_conv2d_maxpool_backward and 16-32 or test_pipelined_input_projection_handover and 3-40 or test_attractor_anchor and 501 or test_model_gradients_golden and convbilstm or test_lstm_seq_wide and 1-1-5 or test_lstm_seq_wide and 2-11-37 or test_training_step_in_stream_groups'
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 --log-file gpurun_out/sanitizer_$TOOL.log python -m pytest tests -m gpu -q -x --timeout 800 -k "$SEL" > gpurun_out/sanitizer_$TOOL.pytest.log 2>&1
  echo "$TOOL exit $?"; tail -3 gpurun_out/sanitizer_$TOOL.pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|hazard" gpurun_out/sanitizer_$TOOL.log | sort | uniq -c | head -20
done
