#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02_tests_e.log 2>&1
echo "gpu tests rc=$?"; tail -3 gpurun_out/r02_tests_e.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_full_e.json 2> gpurun_out/r02_bench_full_e.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_full_e.err
