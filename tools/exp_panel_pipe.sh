#!/bin/bash
# One-call experiment: Cholesky panel variants (0 phase-alternating, 1/2 warp-specialised) x pipelined inverse.
# Every risky combination runs in its own process under a timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/exp_smi.txt 2>&1
timeout 200 python -u tools/gpu_diag.py peak > gpurun_out/exp_peak.log 2>&1; echo "peak rc=$?"
EXP_COMBOS="0:0,0:1" timeout 300 python -u tools/gpu_diag.py exp > gpurun_out/exp_v0.log 2>&1; echo "exp v0 rc=$?"
EXP_COMBOS="2:0,2:1" timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp_v2.log 2>&1; echo "exp v2 rc=$?"
EXP_COMBOS="1:0,1:1" timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp_v1.log 2>&1; echo "exp v1 rc=$?"
timeout 120 python -u tools/gpu_diag.py panel > gpurun_out/exp_panel.log 2>&1; echo "panel rc=$?"
grep -h "VERDICT\|potrf n=\|step \|panel\|chain\|latency\|DFMA\|DMMA" gpurun_out/exp_peak.log gpurun_out/exp_v*.log gpurun_out/exp_panel.log | grep -v "relerr(L)"
read V P <<< "$(python tools/pick_combo.py)"
echo "picked variant=$V pipe=$P"
export MOGP_PANEL_VARIANT=$V MOGP_TRTRI_PIPE=$P
timeout 300 python -u tools/gpu_diag.py lml > gpurun_out/exp_lml_best.log 2>&1; echo "lml rc=$?"
grep -c "FAILED" gpurun_out/exp_lml_best.log; tail -n 40 gpurun_out/exp_lml_best.log
timeout 300 python bench.py --steps 100 --warmup 5 > gpurun_out/exp_bench_best.json 2> gpurun_out/exp_bench_best.err; echo "bench rc=$?"
cat gpurun_out/exp_bench_best.json | cut -c1-600
