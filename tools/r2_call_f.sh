#!/bin/bash
# round 2, GPU call F: suite after the covariance-kernel remap, 2 vs 3 CTAs/SM, ncu --set full of the int8 GEMM / kbuild / grad kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r2f_pytest.log
for mb in 2 3; do
MOGP_COV_MINB=$mb timeout 300 python bench.py --config cfg3 --steps 10 --no-extras --no-cpu-baseline > gpurun_out/r2f_bench_cfg3_minb$mb.json 2> gpurun_out/r2f_bench_cfg3_minb$mb.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_cfg3_minb$mb.json')); print('cfg3 cov_minb=$mb', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
MOGP_COV_MINB=$mb timeout 300 python bench.py --config cfg2 --steps 100 --no-extras --no-cpu-baseline > gpurun_out/r2f_bench_cfg2_minb$mb.json 2> gpurun_out/r2f_bench_cfg2_minb$mb.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_bench_cfg2_minb$mb.json')); print('cfg2 cov_minb=$mb', d['value'], d['ms_per_step'], d['roofline']['stage_ms'])"
done
MOGP_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"i8_gemm_tiles|kbuild_kernel|grad_reduce_kernel|i8_slice_tiled" -c 12 -o gpurun_out/r2f_full_cfg3 python tools/one_step.py --config cfg3 --steps 2 > gpurun_out/r2f_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r2f_full_cfg3.ncu-rep
