#!/bin/bash
# round 2, GPU call L (2 GPUs): bench.py under torchrun, own arm and reference arm
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2l_smi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/r2l_bench_2gpu.json 2> gpurun_out/r2l_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -n 1 gpurun_out/r2l_bench_2gpu.json | cut -c1-700; tail -n 3 gpurun_out/r2l_bench_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > gpurun_out/r2l_bench_ref_2gpu.json 2> gpurun_out/r2l_bench_ref_2gpu.err; echo "bench ref 2gpu rc=$?"; tail -n 1 gpurun_out/r2l_bench_ref_2gpu.json | cut -c1-300
