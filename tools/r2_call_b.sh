#!/bin/bash
# round 2, GPU call B: GPU suite, full bench line (sub-records, reference legs), reference arm, per-stage ncu pass
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r2b_pytest.log
timeout 600 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2b_bench.json; tail -n 5 gpurun_out/r2b_bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2b_bench_reference.json
timeout 300 python bench.py --impl reference-cuda --steps 10 --warmup 2 > gpurun_out/r2b_bench_reference_cuda.json 2> gpurun_out/r2b_bench_reference_cuda.err; echo "refcuda rc=$?"; cut -c1-300 gpurun_out/r2b_bench_reference_cuda.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum
for cfg in cfg2 cfg3; do
  MOGP_GRAPH=0 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2b_step_$cfg.csv python tools/one_step.py --config $cfg --steps 2 > gpurun_out/r2b_ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
  python tools/stage_traffic.py gpurun_out/r2b_step_$cfg.csv $cfg gpurun_out/r2b_stage_traffic.json > gpurun_out/r2b_stage_traffic_$cfg.txt 2>&1; tail -n 3 gpurun_out/r2b_stage_traffic_$cfg.txt
done
