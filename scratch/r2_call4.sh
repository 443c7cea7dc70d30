#!/bin/bash
mkdir -p gpurun_out
PROF_TIME=1 timeout 120 python scratch/prof64.py 2 128 > gpurun_out/r02_call4_times.log 2>&1
PROF_TIME=1 timeout 120 python scratch/prof64.py 2 64 >> gpurun_out/r02_call4_times.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv64k -s 4 -c 1 -o gpurun_out/prof_r02a_conv64k python scratch/prof64.py 1 128 > gpurun_out/r02_call4_ncu.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 12 --csv --log-file gpurun_out/launches_r02a.csv python scratch/prof64.py 2 128 > /dev/null 2>&1
cat gpurun_out/r02_call4_times.log; tail -3 gpurun_out/r02_call4_ncu.log; tail -14 gpurun_out/launches_r02a.csv | cut -c1-200
