#!/bin/bash
# round-2 call 1: prepared A/B variants, fft64 probe, FP32 pipe peak, live-step DRAM traffic (range replay)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_call1_smi.log 2>&1
./scratch/f32x2_bench > gpurun_out/f32x2_bench_r02.log 2>&1
timeout 600 bash scratch/r2_first_call.sh run > gpurun_out/r2_first_call.log 2>&1
timeout 300 ncu --replay-mode range --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum \
    --csv --log-file gpurun_out/live_traffic_r02_base.csv python scratch/prof_range.py 1 > gpurun_out/live_traffic_r02_base.log 2>&1
timeout 300 ncu --replay-mode range --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum \
    --csv --log-file gpurun_out/live_traffic_r02_base4.csv python scratch/prof_range.py 4 > gpurun_out/live_traffic_r02_base4.log 2>&1
tail -3 gpurun_out/f32x2_bench_r02.log gpurun_out/r2_first_call.log gpurun_out/live_traffic_r02_base.csv gpurun_out/live_traffic_r02_base4.csv
