#!/bin/bash
# Round-end evidence (one GPU): per-launch list of the bench command, one `--set full` capture per hot kernel of both
# convolution plans, and LIVE-step range captures (DRAM / L2 bytes and FMA-pipe / issue activity of one steady-state step).
#   gpurun --timeout 1500 -- 'bash scratch/prof_final.sh r02b'
TAG=${1:-r02x}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 12 --warmup 5 --no-cpu --no-extra > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for PLAN in 0 1; do
  NAME=$([ $PLAN = 1 ] && echo block64 || echo partitioned)
  SSB200_BLOCK64=$PLAN timeout 400 ncu --set full --clock-control none --import-source on \
      -k regex:"fwd_rir|mac_bins|mac_ifft|conv64k|spectrogram" -s 16 -c 8 -o gpurun_out/prof_${TAG}_${NAME} \
      python scratch/prof64.py 2 128 > gpurun_out/prof_${TAG}_${NAME}.log 2>&1
  for N in 1 4; do
    SSB200_BLOCK64=$PLAN timeout 300 ncu --replay-mode range --cache-control none --clock-control none \
      --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed \
      --csv --log-file gpurun_out/live_step_${TAG}_${NAME}_${N}.csv python scratch/prof_range.py $N > gpurun_out/live_step_${TAG}_${NAME}_${N}.log 2>&1
  done
done
tail -n 9 gpurun_out/live_step_${TAG}_*_1.csv | cut -c1-220
