#!/bin/bash
# compute-sanitizer over tools/sanitize_driver.py (every kernel once).  Development tool; run on the GPU box.
mkdir -p gpurun_out
for tool in ${1:-memcheck racecheck synccheck initcheck}; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|driver ok" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
  grep -E "^=========\s+at " gpurun_out/sanitize_$tool.log | sed -E 's/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -12
done
