#!/bin/bash
# round 2, GPU call Z: panel variant 3 (split hand-over of the next pivot tile): numerics + timing against variant 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x -k "panel_variants or potrf_variants" > gpurun_out/r2z_pytest_variants.log 2>&1; echo "pytest variants rc=$?"; tail -n 6 gpurun_out/r2z_pytest_variants.log
for v in 3 2; do
MOGP_PANEL_VARIANT=$v DIAG_CFGS=cfg1,cfg2,cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/variant=$v /"
MOGP_PANEL_VARIANT=$v ROWP_COMBOS="1:4096:1:1:0" timeout 120 python tools/gpu_diag.py timeline 2>&1 | grep "^\[timeline" | sed "s/^/variant=$v /"
done
