#!/bin/bash
# Round-end evidence run on the GPU box (one GPU): launch list of the bench command + one `--set full` capture of
# each hot kernel.  Usage: bash scratch/prof_capture.sh <tag>   -> gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep
tag=${1:-r01d}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 32 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 12 --warmup 5 --no-cpu > gpurun_out/bench_under_ncu_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fwd_rir|mac_bins|mac_ifft|spectrogram|logmel" -s 32 -c 9 \
    -f -o gpurun_out/prof_${tag} python scratch/prof_run.py > gpurun_out/prof_${tag}.log 2>&1
ls -la gpurun_out/prof_${tag}.ncu-rep
