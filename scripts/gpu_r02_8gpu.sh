#!/bin/bash
# 8-GPU weak-scaling line of the round (torchrun, NCCL)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 --no-sub-workloads > gpurun_out/r02_bench_cfg2_8gpu.json 2> gpurun_out/r02_bench_cfg2_8gpu.err
echo "rc=$?"; python scripts/show_bench.py gpurun_out/r02_bench_cfg2_8gpu.json
