#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_tests_f.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/r02_tests_f.log
MGV_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-sub-workloads --no-cpu-baseline > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err
echo "bench rc=$?"; grep -E "e2e step|steps," gpurun_out/r02_bench_f.err | tail -14
