#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linear.py -q -m gpu -x > gpurun_out/r02_tests_linear.log 2>&1
echo "linear rc=$?"; tail -3 gpurun_out/r02_tests_linear.log
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_tests_i.log 2>&1
echo "all rc=$?"; tail -3 gpurun_out/r02_tests_i.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-workloads > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err
echo "bench rc=$?"
