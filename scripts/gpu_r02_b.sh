#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "bf16" -s > gpurun_out/r02_tests_bf16.log 2>&1
echo "bf16 rc=$?"; grep -E "worst gradients|passed|failed" gpurun_out/r02_tests_bf16.log | cut -c1-900
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r02_tests_d.log 2>&1
echo "parity rc=$?"; tail -2 gpurun_out/r02_tests_d.log
for w in cfg2 cfg5-k64; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${w}_d.json 2> gpurun_out/r02_bench_${w}_d.err
  echo "bench $w rc=$?"
done
W=cfg5-k8
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis \
   --clock-control none -k regex:sweep_bwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_sweep_bwd_$W python scripts/run_sweep.py $W 4 > gpurun_out/r02_ncu_bwd_$W.log 2>&1
echo "ncu bwd rc=$?"
