#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -s > gpurun_out/r02_tests_cfg.log 2>&1
echo "configs rc=$?"; grep -E "worst grad|passed|failed|decile" gpurun_out/r02_tests_cfg.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 300 python scripts/cprofile_step.py > gpurun_out/r02_cprofile.log 2>&1
echo "cprofile rc=$?"
timeout 300 python scripts/profile_step.py > gpurun_out/r02_profile_step.log 2>&1
echo "profile_step rc=$?"
W=cfg5-k8
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy \
   --clock-control none -k regex:sweep_bwd_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_sweep_bwd_$W python scripts/run_sweep.py $W 4 > gpurun_out/r02_ncu_bwd_$W.log 2>&1
echo "ncu bwd rc=$?"
