#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_call6_topo.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_dropin.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_call6_multirank_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r02_call6_bench_2gpu.json 2> gpurun_out/r02_call6_bench_2gpu.err
tail -n 8 gpurun_out/r02_call6_multirank_tests.log; tail -n 3 gpurun_out/r02_call6_bench_2gpu.err; head -c 600 gpurun_out/r02_call6_bench_2gpu.json
