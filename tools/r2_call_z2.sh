#!/bin/bash
# round 2, GPU call Z2: what bounds the panel pitch?  bare chain (no trailing updates; wrong numbers, timing only) for variants 2 / 3
mkdir -p gpurun_out
for v in 3 2; do
for sb in 1 0; do
MOGP_SKIP_BULK=$sb MOGP_PANEL_VARIANT=$v ROWP_COMBOS="1:4096:1:1:0,0:4096:1:1:0" timeout 120 python tools/gpu_diag.py timeline 2>&1 | grep "^\[timeline\|spans:" | sed "s/^/variant=$v skip_bulk=$sb /"
done; done
MOGP_PANEL_VARIANT=3 timeout 200 python tools/gpu_diag.py spans 2>&1 | grep "spans n=" | head -8
