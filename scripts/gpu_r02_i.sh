#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_readout.py -q -m gpu > gpurun_out/r02_tests_readout.log 2>&1
echo "readout rc=$?"; tail -2 gpurun_out/r02_tests_readout.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub-workloads > gpurun_out/r02_bench_k.json 2> gpurun_out/r02_bench_k.err
echo "bench rc=$?"
