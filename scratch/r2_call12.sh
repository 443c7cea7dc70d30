#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_r02_8gpu.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/bench_r02c_8gpu.json 2> gpurun_out/bench_r02c_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 8 --steps 100 --warmup 10 --no-numa > gpurun_out/bench_r02c_8gpu_nonuma.json 2> gpurun_out/bench_r02c_8gpu_nonuma.err
MR_ENVS=128 timeout 600 python -m pytest tests/test_gpu_multirank.py -q -m gpu 2>&1 | tail -4 > gpurun_out/gpu_tests_r02c_8gpu_multirank.log
tail -n 3 gpurun_out/bench_r02c_8gpu.err; head -c 300 gpurun_out/bench_r02c_8gpu.json; echo; cat gpurun_out/gpu_tests_r02c_8gpu_multirank.log
