#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "boundary or replay or readout or edge_split or train_step" 2>&1 | tail -3 | tee gpurun_out/k_tests.log
MGV_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-sub-workloads > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/k_bench.json 2>&1 | tail -30
grep "e2e step" gpurun_out/k_bench.err | tail -6
timeout 300 python scripts/cprofile_step.py cfg2 > gpurun_out/k_cprofile.txt 2>&1
head -70 gpurun_out/k_cprofile.txt | tail -62
