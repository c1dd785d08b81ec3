#!/bin/bash
# compute-sanitizer over every kernel on small shapes.  One tool per process, each under a timeout.
mkdir -p gpurun_out
timeout 120 python scripts/sanitize_case.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain exit $?"; tail -1 gpurun_out/sanitize_plain.log
for tool in memcheck synccheck racecheck; do
  timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize case ok|Error|hazard" gpurun_out/sanitize_$tool.log | head -8
done
