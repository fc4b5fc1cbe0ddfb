#!/bin/bash
# round 2, GPU call Y: replicas stacked on one GPU (cfg2 = the cfg5 restart workload)
mkdir -p gpurun_out
timeout 900 python tools/replica_sweep.py cfg2 2,3,4,6,8 > gpurun_out/r2y_replica_sweep.log 2>&1; echo "rc=$?"; cat gpurun_out/r2y_replica_sweep.log | tail -8
