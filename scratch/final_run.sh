#!/bin/bash
# Round-end validation on one GPU: GPU tests, smoke, both bench arms at the driver's settings and at the long setting.
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/gpu_tests_${TAG}.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_${TAG}.log 2>&1
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2> gpurun_out/bench_${TAG}_reference_arm.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_driver_setting.json 2> gpurun_out/bench_${TAG}_driver_setting.err
timeout 900 python bench.py --steps 300 --warmup 30 > gpurun_out/bench_${TAG}_final.json 2> gpurun_out/bench_${TAG}_final.err
tail -n 3 gpurun_out/gpu_tests_${TAG}.log gpurun_out/smoke_${TAG}.log; tail -n 2 gpurun_out/bench_${TAG}_final.err
head -c 400 gpurun_out/bench_${TAG}_driver_setting.json; echo; head -c 400 gpurun_out/bench_${TAG}_final.json; echo; head -c 300 gpurun_out/bench_${TAG}_reference_arm.json
