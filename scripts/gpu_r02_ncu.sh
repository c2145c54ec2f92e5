#!/bin/bash
# ncu --set full captures of the two level-sweep kernels (one launch each, after warm-up), source-level counters included
mkdir -p gpurun_out
W=${1:-cfg5-k8}
python scripts/run_sweep.py $W 4 > gpurun_out/r02_run_sweep_$W.log 2>&1 || exit 1
for k in fwd bwd; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_${k}_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_sweep_${k}_$W \
      python scripts/run_sweep.py $W 4 > gpurun_out/r02_ncu_${k}_$W.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
