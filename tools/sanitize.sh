#!/bin/bash
# compute-sanitizer over a small slice of the GPU tests (memcheck + racecheck + synccheck)
mkdir -p gpurun_out
SEL='test_stft_vs_oracle and 777 or test_istft_vs_oracle and 33 or test_center or test_mix_features or test_linear and 37 or test_lstm_seq and 2-3-12 or test_attractor_anchor and 3-2-9 or test_attractor_truth and 3-2-9 or test_mask_cmul and 3-2-9 or test_pit_mse and 4-2-7 or test_head_backward and 3-2-9 or test_lstm_layer_backward and 3-9-40 or test_clip_adam'
for TOOL in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $TOOL --print-limit 20 --log-file gpurun_out/sanitizer_$TOOL.log python -m pytest tests -m gpu -q -x --timeout 800 -k "$SEL" > gpurun_out/sanitizer_$TOOL.pytest.log 2>&1
  echo "$TOOL exit $?"; tail -3 gpurun_out/sanitizer_$TOOL.pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Race reported|Invalid|hazard" gpurun_out/sanitizer_$TOOL.log | sort | uniq -c | head -20
done
