#!/bin/bash
# Steady-state DRAM traffic per kernel: no cache flush between kernels (--cache-control none), so L2 carries over as in
# the live step (kernels are serialised by the profiler, so the two sub-batch chains do not overlap here).
mkdir -p gpurun_out
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum \
    -k regex:"fwd_rir|mac_bins|mac_ifft|spectrogram" -s 64 -c 16 --csv --log-file gpurun_out/dram_traffic_${1:-x}.csv \
    python scratch/prof_steady.py ${2:-2} ${3:-1} > gpurun_out/dram_traffic_${1:-x}.log 2>&1
