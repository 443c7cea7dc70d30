#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/r02_call7_tests.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scratch/sanitize_run.py > gpurun_out/sanitizer_r02_$tool.log 2>&1
  tail -n 3 gpurun_out/sanitizer_r02_$tool.log
done
tail -n 5 gpurun_out/r02_call7_tests.log
