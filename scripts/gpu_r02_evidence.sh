#!/bin/bash
# round-2 evidence: ncu captures of the level-sweep kernels, launch list of a step, compute-sanitizer runs
mkdir -p gpurun_out
for W in cfg2 cfg5-k64; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_fwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_sweep_fwd_tc_$W \
      python scripts/run_sweep.py $W 4 > gpurun_out/r02_ncu_fwd_$W.log 2>&1
  echo "ncu fwd $W rc=$?"
  timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section WarpStateStats --section SchedulerStats \
      --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --clock-control none -k regex:sweep_bwd_tc_kernel -s 2 -c 1 -f \
      -o gpurun_out/r02_sweep_bwd_tc_$W python scripts/run_sweep.py $W 4 > gpurun_out/r02_ncu_bwd_$W.log 2>&1
  echo "ncu bwd $W rc=$?"
done
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python scripts/run_small.py mig 6 300 > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; tail -3 gpurun_out/r02_sanitizer_$tool.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-workloads > gpurun_out/r02_ncu_launch.log 2>&1
echo "launch list rc=$?"
