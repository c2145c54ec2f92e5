#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_readout.py -x -q -m gpu > gpurun_out/r02_tests_readout.log 2>&1
echo "readout rc=$?"; tail -5 gpurun_out/r02_tests_readout.log
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02_tests_g.log 2>&1
echo "all gpu tests rc=$?"; tail -3 gpurun_out/r02_tests_g.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sub-workloads > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err
echo "bench rc=$?"
