#!/bin/bash
# round 2, GPU call I: wide (128 x 128, two-pass) int8 GEMM variant against the one-pass kernel
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py i8p > gpurun_out/r2i_i8p.log 2>&1; echo "i8p rc=$?"; grep -E "^---|8192x 8192x 8192|4096x 4096x 4096|8192x 8192x 1024" gpurun_out/r2i_i8p.log
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -k "int8" > gpurun_out/r2i_pytest_i8.log 2>&1; echo "pytest i8 rc=$?"; tail -n 5 gpurun_out/r2i_pytest_i8.log
for w in 0 1; do
MOGP_I8_WIDE=$w timeout 300 python bench.py --config cfg3 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2i_bench_cfg3_w$w.json 2> gpurun_out/r2i_bench_cfg3_w$w.err; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_cfg3_w$w.json')); print('cfg3 wide=$w', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
MOGP_I8_WIDE=$w timeout 300 python bench.py --config cfg4 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2i_bench_cfg4_w$w.json 2> gpurun_out/r2i_bench_cfg4_w$w.err; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_cfg4_w$w.json')); print('cfg4 wide=$w', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
done
MOGP_I8_WIDE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_dropin.py -m gpu -q -k "cfg3 or cfg4" > gpurun_out/r2i_pytest_big.log 2>&1; echo "pytest big (wide on) rc=$?"; tail -n 5 gpurun_out/r2i_pytest_big.log
