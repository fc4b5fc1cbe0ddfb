#!/bin/bash
# ncu --set full captures of the two top kernels of the cfg2 step (eager launches: MOGP_GRAPH=0).
mkdir -p gpurun_out
MOGP_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:potrf_panel_ws_kernel -s 40 -c 2 -f -o gpurun_out/r01_full_panel_ws python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_panel.log 2>&1; echo "ncu panel rc=$?"
MOGP_GRAPH=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_f64_kernel -s 230 -c 6 -f -o gpurun_out/r01_full_gemm_cfg2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_gemm.log 2>&1; echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
