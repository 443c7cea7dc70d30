#!/bin/bash
# Round-end validation on the GPU box: full GPU test-suite, smoke, bench (both arms), sanitizer, ncu evidence.
tag=${1:-r01d}
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -q -m gpu 2>&1 | tail -6) > gpurun_out/final_tests_${tag}.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke_${tag}.log 2>&1
timeout 200 python bench.py --steps 300 --warmup 30 > gpurun_out/bench_${tag}_final.json 2> gpurun_out/bench_${tag}_final.err
timeout 100 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_${tag}_final_reference_arm.json 2>/dev/null
timeout 150 bash scratch/prof_capture.sh ${tag}
timeout 150 compute-sanitizer --tool memcheck python scratch/sanitize_run.py 2>&1 | grep -E "=========|done" | tail -8 > gpurun_out/sanitizer_${tag}_memcheck.log
cat gpurun_out/final_tests_${tag}.log gpurun_out/final_smoke_${tag}.log gpurun_out/sanitizer_${tag}_memcheck.log
cut -c1-400 gpurun_out/bench_${tag}_final.json
