#!/bin/bash
# compute-sanitizer passes over a small end-to-end scenario of every model family (smoke() = monod, 20 000 particles,
# 500 compartments, divisions, exits, compaction; tools/sanitize_case.py = the other models, 0D, eager ages, forced
# compaction, export, checkpoint).  Logs -> gpurun_out/sanitize_<tool>.log
T=${TAG:-sanitize}
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no"
  [ $tool = racecheck ] && extra="--racecheck-report analysis"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool $extra --error-exitcode 7 --print-limit 2000 python tools/sanitize_case.py > gpurun_out/${T}_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/${T}_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|case ok|all cases ok" gpurun_out/${T}_$tool.log | sort | uniq -c | head -20
  grep -oE "bmc_[a-z_]+\.cuh:[0-9]+" gpurun_out/${T}_$tool.log | sort | uniq -c | sort -rn | head -20
done
