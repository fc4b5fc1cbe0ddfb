#!/bin/bash
# round 2, GPU call R: two-level row-wise pipeline (groups / super-groups / taper): A/B, timeline, knob tests
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py rowp > gpurun_out/r2r_rowp.log 2>&1; echo "rowp rc=$?"; grep -v "^\[.*trtri" gpurun_out/r2r_rowp.log | tail -n 45
grep "trtri" gpurun_out/r2r_rowp.log | awk '{print $NF, $(NF-2)}' | sort -g | tail -n 3
timeout 300 python tools/gpu_diag.py timeline > gpurun_out/r2r_timeline.log 2>&1; echo "timeline rc=$?"; grep "^\[timeline" gpurun_out/r2r_timeline.log
timeout 900 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x -k "rowwise" > gpurun_out/r2r_pytest_rowwise.log 2>&1; echo "pytest rowwise rc=$?"; tail -n 5 gpurun_out/r2r_pytest_rowwise.log
