#!/bin/bash
mkdir -p gpurun_out
SEL='test_lstm_layer_backward and 600 or test_lstm_layer_backward and 416 or test_lstm_seq_wide and 1-1-5'
for TOOL in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 --log-file gpurun_out/sanitizer_wide_$TOOL.log python -m pytest tests -m gpu -q -x --timeout 1100 -k "$SEL" > gpurun_out/sanitizer_wide_$TOOL.pytest.log 2>&1
  echo "$TOOL exit $?"; tail -2 gpurun_out/sanitizer_wide_$TOOL.pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitizer_wide_$TOOL.log
done
