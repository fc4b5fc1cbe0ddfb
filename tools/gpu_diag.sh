#!/bin/bash
# Bring-up diagnostics: every section in its own process, logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/diag_smi.txt 2>&1
for s in "$@"; do
  timeout 600 python tools/gpu_diag.py $s > gpurun_out/diag_$s.log 2>&1
  echo "section $s rc=$?"
  tail -n 60 gpurun_out/diag_$s.log
done
