#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
  if [ $i -gt 3 ]; then export MGV_BENCH_NO_SAMPLER=1; fi
  MGV_BENCH_VERBOSE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-sub-workloads --no-cpu-baseline > gpurun_out/r02_e2e_$i.json 2> gpurun_out/r02_e2e_$i.err
  python - <<PY
import json,re
d=json.load(open("gpurun_out/r02_e2e_$i.json"))
t=[float(x) for x in re.findall(r"train_step host ([0-9.]+)", open("gpurun_out/r02_e2e_$i.err").read())]
print("run $i (nosampler=%s): resident %.2f ms  e2e %.2f ms  worst host step %.1f ms at %d" % ("$MGV_BENCH_NO_SAMPLER", d["ms_per_step"], d["e2e"]["ms_per_step"], max(t), t.index(max(t))))
PY
done
