#!/bin/bash
# Round validation on the GPU box: full GPU test suite, smoke(), bench lines (cfg2 default, cfg3), reference arm,
# ncu launch list of one bench step.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/val_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/val_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/val_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/val_smoke.log
timeout 400 python bench.py > gpurun_out/val_bench_cfg2.json 2> gpurun_out/val_bench_cfg2.err; echo "bench cfg2 rc=$?"; cut -c1-300 gpurun_out/val_bench_cfg2.json
timeout 400 python bench.py --config cfg3 --steps 20 --no-cpu-baseline > gpurun_out/val_bench_cfg3.json 2> gpurun_out/val_bench_cfg3.err; echo "bench cfg3 rc=$?"; cut -c1-300 gpurun_out/val_bench_cfg3.json
if [ "$1" == "full" ]; then
  timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/val_bench_reference.json 2> gpurun_out/val_bench_reference.err; echo "bench ref rc=$?"; cut -c1-300 gpurun_out/val_bench_reference.json
  MOGP_GRAPH=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/val_launches_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/val_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
  python tools/launch_summary.py gpurun_out/val_launches_cfg2.csv > gpurun_out/val_launches_cfg2_summary.txt 2>&1; head -n 25 gpurun_out/val_launches_cfg2_summary.txt
fi
