#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" EXP_COMBOS="2:1" EXP_SKIP_NUM=1 timeout 200 python -u tools/gpu_diag.py exp > gpurun_out/exp14_$name.log 2>&1; echo "$name rc=$?"
  grep -h "VERDICT\|step " gpurun_out/exp14_$name.log | sed "s/^/[$name] /" | cut -c1-330
}
run pdl1 EXP_PDL=1 DIAG_CFGS=cfg2,cfg1
run pdl1_nofence EXP_PDL=1 EXP_NOFENCE=1 DIAG_CFGS=cfg2,cfg1
run pdl1_graphall EXP_PDL=1 EXP_GRAPH_MAX_NP=8192 DIAG_CFGS=cfg4,cfg3
run pdl0_graphall EXP_PDL=0 EXP_GRAPH_MAX_NP=8192 DIAG_CFGS=cfg4,cfg3
run pdl2_eager EXP_PDL=2 DIAG_CFGS=cfg4
