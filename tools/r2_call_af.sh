#!/bin/bash
# round 2, GPU call AF: (a) recursive scheme at N = 2048 with 1024-row leaves, (b) wide int8 kernel for every product
mkdir -p gpurun_out
MOGP_RCHOL_MIN_NP=2048 MOGP_RCHOL_LEAF=1024 DIAG_CFGS=cfg2,cfg4 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/leaf1024 /"
DIAG_CFGS=cfg2,cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/default /"
MOGP_I8_WIDE=3 DIAG_CFGS=cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/wide3 /"
MOGP_I8_WIDE=0 DIAG_CFGS=cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/wide0 /"
MOGP_I8_SLICES=8 DIAG_CFGS=cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/slices8 /"
