#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 400 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|bands equal|done|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
