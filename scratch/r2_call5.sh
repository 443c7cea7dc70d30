#!/bin/bash
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/r02_call5_smoke.log 2>&1
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_call5_tests.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_call5_bench.json 2> gpurun_out/r02_call5_bench.err
tail -n 6 gpurun_out/r02_call5_smoke.log gpurun_out/r02_call5_tests.log; tail -n 5 gpurun_out/r02_call5_bench.err; head -c 1500 gpurun_out/r02_call5_bench.json
