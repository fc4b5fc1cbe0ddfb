#!/bin/bash
mkdir -p gpurun_out
for v in 4 2; do
MOGP_PANEL_VARIANT=$v MOGP_TRTRI_PIPE=0 MOGP_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/exp8_launches_v$v.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/exp8_ncu_v$v.log 2>&1; echo "ncu v$v rc=$?"
python tools/launch_summary.py gpurun_out/exp8_launches_v$v.csv | head -6
python tools/launch_summary.py gpurun_out/exp8_launches_v$v.csv --detail | grep "potrf_panel" | awk '{print $NF, $(NF-1)}' | tr '\n' ' ' | cut -c1-700; echo
done
