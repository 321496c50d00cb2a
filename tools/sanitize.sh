#!/bin/bash
# compute-sanitizer pass over the structured kernels (run on a GPU box: gpurun -- tools/sanitize.sh).
# memcheck + synccheck on all cases, racecheck (shared-memory hazards: the mbarrier rings, the mid planes of the
# single-pass two-colour sweeps) on the first two.  Summaries go to gpurun_out/sanitize_*.log.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck; do
  timeout 900 $CS --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_driver.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_$tool.log
  grep -E "ERROR SUMMARY|sanitize_driver: done" gpurun_out/sanitize_$tool.log
done
timeout 1500 $CS --tool racecheck --racecheck-report all --error-exitcode 9 --print-limit 20 python tools/sanitize_driver.py 2 > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_racecheck.log
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|sanitize_driver: done" gpurun_out/sanitize_racecheck.log
