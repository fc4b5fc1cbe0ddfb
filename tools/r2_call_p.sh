#!/bin/bash
# round 2, GPU call P: row-wise pipelined inverse + progressive K^-1 (A/B against the block-doubling pipeline), knob tests
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py rowp > gpurun_out/r2p_rowp.log 2>&1; echo "rowp rc=$?"; grep -v "^\[.*trtri" gpurun_out/r2p_rowp.log | tail -n 45
grep "trtri" gpurun_out/r2p_rowp.log | awk '{print $NF, $(NF-2)}' | sort -g | tail -n 3
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x -k "rowwise" > gpurun_out/r2p_pytest_rowwise.log 2>&1; echo "pytest rowwise rc=$?"; tail -n 5 gpurun_out/r2p_pytest_rowwise.log
