#!/bin/bash
# round 2, GPU call G: TS-form int8 GEMM (A planes in tensor memory) bring-up, concurrent replicas, full bench line
mkdir -p gpurun_out
I8_TS=1 timeout 200 python tools/gpu_diag.py i8p > gpurun_out/r2g_i8p_ts1.log 2>&1; echo "i8p ts=1 rc=$?"; tail -n 9 gpurun_out/r2g_i8p_ts1.log
I8_TS=0 timeout 200 python tools/gpu_diag.py i8p > gpurun_out/r2g_i8p_ts0.log 2>&1; echo "i8p ts=0 rc=$?"; tail -n 5 gpurun_out/r2g_i8p_ts0.log
timeout 900 python -m pytest tests/test_gpu_knobs.py tests/test_gpu_model.py -m gpu -q -k "int8 or concurrent" > gpurun_out/r2g_pytest_sel.log 2>&1; echo "pytest sel rc=$?"; tail -n 8 gpurun_out/r2g_pytest_sel.log
for ts in 0 1; do
MOGP_I8_TS=$ts timeout 300 python bench.py --config cfg3 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2g_bench_cfg3_ts$ts.json 2> gpurun_out/r2g_bench_cfg3_ts$ts.err; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_cfg3_ts$ts.json')); print('cfg3 ts=$ts', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
done
timeout 900 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2g_bench.json')); print(d['value'], d['e2e']['value'], d['e2e'].get('concurrent_replicas'))"; tail -n 3 gpurun_out/r2g_bench.err
