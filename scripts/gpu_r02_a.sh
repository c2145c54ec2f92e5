#!/bin/bash
# round-2 GPU check of the multi-stream level sweep: parity tests, bench lines, phase traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu > gpurun_out/r02_tests_c.log 2>&1
echo "tests rc=$?"
tail -3 gpurun_out/r02_tests_c.log
for w in cfg2 cfg5-k8 cfg5-k64; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${w}_c.json 2> gpurun_out/r02_bench_${w}_c.err
  echo "bench $w rc=$?"
done
export MGV_B200_LIB=/root/repo/build/_trace/libmgv_b200.so
for w in cfg2 cfg5-k64; do
  for d in fwd bwd; do
    echo "== $w $d"
    timeout 300 python scripts/trace_sweep_tc.py $w $d 2>&1 | tail -24
  done
done > gpurun_out/r02_trace_c.log 2>&1
echo done
