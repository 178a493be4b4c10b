#!/bin/bash
# usage (GPU box): tools/sanitize.sh [out_dir]   -> compute-sanitizer memcheck / racecheck / synccheck over the smoke runs
cd "$(dirname "$0")/.."
out=${1:-gpurun_out}
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_smoke.py > $out/r02_sanitizer_$tool.log 2>&1
  echo "$tool: exit $?  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/r02_sanitizer_$tool.log | tail -1)"
done
