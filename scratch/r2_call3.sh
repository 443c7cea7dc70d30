#!/bin/bash
# round-2 call 3: first contact of the single-block cluster kernel
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02_call3_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_call3_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r02_call3_parity.log
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r02_call3_tests.log
SSB200_BLOCK64=0 timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/r02_call3_bench_legacy.json 2> gpurun_out/r02_call3_bench_legacy.err
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu > gpurun_out/r02_call3_bench_block64.json 2> gpurun_out/r02_call3_bench_block64.err
timeout 300 ncu --replay-mode range --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
    --csv --log-file gpurun_out/live_traffic_r02_block64.csv python scratch/prof_range.py 1 > gpurun_out/live_traffic_r02_block64.log 2>&1
tail -n 4 gpurun_out/r02_call3_smoke.log gpurun_out/r02_call3_parity.log gpurun_out/r02_call3_tests.log
head -c 300 gpurun_out/r02_call3_bench_legacy.json; echo; head -c 300 gpurun_out/r02_call3_bench_block64.json; echo
tail -n 4 gpurun_out/live_traffic_r02_block64.csv
