#!/bin/bash
# End-of-round validation on the GPU box: full GPU suite, smoke(), the default bench line + the two reference arms, the
# per-stage ncu pass (eager launches, no cache flush) of cfg2 / cfg3 and one ncu --set full capture of the hot kernels.
mkdir -p gpurun_out
P=${1:-val}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/${P}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${P}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${P}_smoke.log
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err; echo "bench ref rc=$?"; cut -c1-200 gpurun_out/${P}_bench_reference.json
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${P}_bench.json
timeout 300 python bench.py --impl reference-cuda --steps 10 --warmup 2 > gpurun_out/${P}_bench_reference_cuda.json 2> gpurun_out/${P}_bench_reference_cuda.err; echo "bench ref-cuda rc=$?"; cut -c1-200 gpurun_out/${P}_bench_reference_cuda.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum
for cfg in cfg2 cfg3; do
  MOGP_GRAPH=0 timeout 600 ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/${P}_step_$cfg.csv python tools/one_step.py --config $cfg --steps 3 > gpurun_out/${P}_ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
  python tools/stage_traffic.py gpurun_out/${P}_step_$cfg.csv $cfg gpurun_out/${P}_stage_traffic.json > gpurun_out/${P}_stage_traffic_$cfg.txt 2>&1
done
MOGP_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"i8_gemm_wide|i8_gemm_tiles|potrf_panel_ws" -s 140 -c 14 -o gpurun_out/${P}_full_cfg3 python tools/one_step.py --config cfg3 --steps 2 > gpurun_out/${P}_ncu_full.log 2>&1; echo "ncu full rc=$?"
