#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/k_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-sub-workloads --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/k_bench.json 2>&1 | tail -3
