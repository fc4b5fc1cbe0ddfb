#!/bin/bash
# round 2, GPU call AC: panel variant 3 with two accumulator chains per tile in the pre-accumulation
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x -k "panel_variants or potrf_variants" > gpurun_out/r2ac_pytest_variants.log 2>&1; echo "pytest variants rc=$?"; tail -n 4 gpurun_out/r2ac_pytest_variants.log
for v in 3 2 3 2; do
MOGP_PANEL_VARIANT=$v DIAG_CFGS=cfg2,cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/variant=$v /"
MOGP_PANEL_VARIANT=$v ROWP_COMBOS="1:4096:1:1:0" timeout 120 python tools/gpu_diag.py timeline 2>&1 | grep "^\[timeline" | sed "s/^/variant=$v /"
done
MOGP_PANEL_VARIANT=3 timeout 200 python tools/gpu_diag.py spans 2>&1 | grep "spans n=" | head -2
