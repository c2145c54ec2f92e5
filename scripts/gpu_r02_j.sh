#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/r02_tests_k.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/r02_tests_k.log
for i in 1 2; do
MGV_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub-workloads > gpurun_out/r02_bench_l$i.json 2> gpurun_out/r02_bench_l$i.err
echo "bench rc=$?"; grep cudaMalloc gpurun_out/r02_bench_l$i.err
done
