#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 250 python -u tools/gpu_diag.py exp > gpurun_out/exp15_$name.log 2>&1; echo "$name rc=$?"
  grep -h "VERDICT\|step \|potrf n=8192\|trtri n=" gpurun_out/exp15_$name.log | sed "s/^/[$name] /" | cut -c1-330; tail -2 gpurun_out/exp15_$name.log | cut -c1-300
}
run pipe1 EXP_PDL=1 EXP_GRAPH_MAX_NP=1000000 EXP_COMBOS="2:1" DIAG_CFGS=cfg3,cfg4
run pipe0 EXP_PDL=1 EXP_GRAPH_MAX_NP=1000000 EXP_COMBOS="2:0" DIAG_CFGS=cfg3
