#!/bin/bash
# round-2 final measurements on one B200 (REF=1 also runs the reference arm: ~3 minutes of host time)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_gputests_final.log 2>&1
echo "gpu tests rc=$?"; tail -2 gpurun_out/r02_gputests_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke_final.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg2.json 2> gpurun_out/r02_bench_cfg2.err
echo "bench rc=$?"
if [ -n "$REF" ]; then
  timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_cfg2_reference_arm.json 2> gpurun_out/r02_bench_cfg2_reference_arm.err
  echo "reference arm rc=$?"
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-workloads > gpurun_out/r02_ncu_launch.log 2>&1
echo "launch list rc=$?"
python scripts/show_bench.py gpurun_out/r02_bench_cfg2.json
