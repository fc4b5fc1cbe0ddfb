#!/bin/bash
# round 2, GPU call J: width selection of the int8 GEMM per use, Python-loop host profile
mkdir -p gpurun_out
for w in 1 2; do
MOGP_I8_WIDE=$w timeout 300 python bench.py --config cfg3 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2j_bench_cfg3_w$w.json 2> gpurun_out/r2j_bench_cfg3_w$w.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_cfg3_w$w.json')); print('cfg3 wide=$w', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
done
MOGP_I8_WIDE=1 timeout 300 python bench.py --config cfg4 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2j_bench_cfg4_w1.json 2> gpurun_out/r2j_bench_cfg4_w1.err; python -c "
import json; d=json.load(open('gpurun_out/r2j_bench_cfg4_w1.json')); print('cfg4 wide=1', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
timeout 300 python tools/profile_python_loop.py > gpurun_out/r2j_python_loop.txt 2>&1; echo "profile rc=$?"; cat gpurun_out/r2j_python_loop.txt | grep -v Warning | tail -n 12
timeout 900 python -m pytest tests/test_gpu_knobs.py tests/test_gpu_parity.py -m gpu -q -k "int8 or cfg3 or cfg4" > gpurun_out/r2j_pytest_sel.log 2>&1; echo "pytest sel rc=$?"; tail -n 4 gpurun_out/r2j_pytest_sel.log
