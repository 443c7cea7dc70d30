#!/bin/bash
mkdir -p gpurun_out
for G in 2 4; do
  SSB200_C64_GROUPS=$G timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02_call8_parity_g$G.log
  SSB200_C64_GROUPS=$G SSB200_BLOCK64=1 PROF_TIME=1 timeout 120 python scratch/prof64.py 2 128 > gpurun_out/r02_call8_times_g$G.log 2>&1
  tail -n 2 gpurun_out/r02_call8_parity_g$G.log; cat gpurun_out/r02_call8_times_g$G.log
done
SSB200_BLOCK64=0 PROF_TIME=1 timeout 120 python scratch/prof64.py 2 128 > gpurun_out/r02_call8_times_partitioned.log 2>&1; cat gpurun_out/r02_call8_times_partitioned.log
