#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -m gpu -x > gpurun_out/r02_tests_h.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/r02_tests_h.log
for w in cfg2 cfg5-k64; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-sub-workloads > gpurun_out/r02_bench_${w}_h.json 2> gpurun_out/r02_bench_${w}_h.err
  echo "bench $w rc=$?"
done
