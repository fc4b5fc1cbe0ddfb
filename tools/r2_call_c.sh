#!/bin/bash
# round 2, GPU call C: int8 tensor-pipe GEMM (product kernel) bring-up + parity at N >= 4096, traffic pass without cache flush
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py i8p > gpurun_out/r2c_i8p.log 2>&1; echo "i8p rc=$?"; tail -n 16 gpurun_out/r2c_i8p.log
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -k "int8" > gpurun_out/r2c_pytest_i8.log 2>&1; echo "pytest i8 rc=$?"; tail -n 8 gpurun_out/r2c_pytest_i8.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r2c_pytest.log
timeout 300 python bench.py --config cfg3 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err; echo "bench cfg3 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_cfg3.json')); print(d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
MOGP_I8_MIN_NP=0 timeout 300 python bench.py --config cfg3 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2c_bench_cfg3_dmma.json 2> gpurun_out/r2c_bench_cfg3_dmma.err; python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_cfg3_dmma.json')); print('dmma', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum
for cfg in cfg2 cfg3; do
  MOGP_GRAPH=0 timeout 600 ncu --metrics $M --clock-control none --cache-control none --csv --log-file gpurun_out/r2c_step_$cfg.csv python tools/one_step.py --config $cfg --steps 3 > gpurun_out/r2c_ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
  python tools/stage_traffic.py gpurun_out/r2c_step_$cfg.csv $cfg gpurun_out/r2c_stage_traffic.json > gpurun_out/r2c_stage_traffic_$cfg.txt 2>&1; tail -n 3 gpurun_out/r2c_stage_traffic_$cfg.txt
done
