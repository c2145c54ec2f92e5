#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "bf16 or cfg4 or struct" 2>&1 | tail -15 | tee gpurun_out/k_tests.log
timeout 600 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline --no-sub-workloads > gpurun_out/k_bench_cfg4.json 2> gpurun_out/k_bench_cfg4.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/k_bench_cfg4.json 2>&1 | tail -3
