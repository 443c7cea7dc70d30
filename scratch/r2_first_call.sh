#!/bin/bash
# Round-2 opener: build the prepared, not-yet-measured variants HERE (no GPU needed), then A/B them on the box:
#   bash scratch/r2_first_call.sh build      # in the build container
#   gpurun --timeout 300 -- 'bash scratch/r2_first_call.sh run'
set -e
FL="-shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17"
SRC=soundspaces_b200/csrc/ssb200.cu
case "$1" in
build)
    mkdir -p scratch/ab
    nvcc $FL -o scratch/ab/lib_base.so $SRC &
    nvcc $FL -DMACB_PREFETCH=3 -o scratch/ab/lib_pf3.so $SRC &
    nvcc $FL -DMACB_PREFETCH=4 -o scratch/ab/lib_pf4.so $SRC &
    nvcc $FL -DMACB_PREFETCH=6 -o scratch/ab/lib_pf6.so $SRC &
    nvcc $FL -DMACB_PREFETCH=4 -DMACB_MIN_BLOCKS=4 -o scratch/ab/lib_pf4_mc4.so $SRC &
    nvcc $FL -DIFFT_TW_GLOBAL=1 -o scratch/ab/lib_twg.so $SRC &
    nvcc $FL -DIFFT_TW_GLOBAL=1 -DMACB_PREFETCH=4 -o scratch/ab/lib_twg_pf4.so $SRC &
    nvcc $FL -o scratch/libfft64_probe.so scratch/fft64_probe.cu &
    wait; ls -la scratch/ab scratch/libfft64_probe.so ;;
run)
    mkdir -p gpurun_out
    python scratch/ab_libs.py scratch/ab/lib_base.so scratch/ab/lib_pf3.so scratch/ab/lib_pf4.so scratch/ab/lib_pf6.so \
        scratch/ab/lib_pf4_mc4.so scratch/ab/lib_twg.so scratch/ab/lib_twg_pf4.so scratch/ab/lib_base.so 2>&1 | grep -E "^AB|rror" | tee gpurun_out/ab_r02_prepared.log
    python scratch/fft64_probe.py 2>&1 | tee gpurun_out/fft64_probe_r02.log
    # parity of the winner: SSB200_LIB=$PWD/scratch/ab/lib_pf4.so python -m pytest tests/test_gpu_parity.py -q -m gpu
    ;;
*) echo "usage: $0 build|run" ;;
esac
