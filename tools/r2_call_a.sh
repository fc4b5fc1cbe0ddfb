#!/bin/bash
# round 2, GPU call A: full GPU suite (incl. the new reference drop-in tests), int8 tcgen05 GEMM bring-up, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 30 gpurun_out/r2a_pytest.log
I8_VARIANT=0 timeout 200 python tools/gpu_diag.py i8 > gpurun_out/r2a_i8_v0.log 2>&1; echo "i8 v0 rc=$?"; tail -n 16 gpurun_out/r2a_i8_v0.log
I8_VARIANT=1 timeout 200 python tools/gpu_diag.py i8 > gpurun_out/r2a_i8_v1.log 2>&1; echo "i8 v1 rc=$?"; tail -n 16 gpurun_out/r2a_i8_v1.log
timeout 400 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2a_bench.json
