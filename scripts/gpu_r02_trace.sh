#!/bin/bash
mkdir -p gpurun_out
export MGV_B200_LIB=/root/repo/build/_trace/libmgv_b200.so
for w in cfg2 cfg5-k64; do
  for d in bwd; do
    echo "== $w $d"
    timeout 300 python scripts/trace_sweep_tc.py $w $d 2>&1 | tail -30
  done
done > gpurun_out/r02_trace_d.log 2>&1
echo done
