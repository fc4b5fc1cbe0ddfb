#!/bin/bash
# round 2, GPU call AM: diagonal inverse + X row product fused into one launch (MOGP_XFUSE), A/B on one box + numerics
mkdir -p gpurun_out
MOGP_XFUSE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/r2am_pytest.log 2>&1; echo "pytest xfuse rc=$?"; tail -n 2 gpurun_out/r2am_pytest.log
MOGP_XFUSE=1 timeout 600 python -m pytest tests/test_gpu_knobs.py -m gpu -q -x -k "rowwise or recursive or panel_variants" > gpurun_out/r2am_pytest_knobs.log 2>&1; echo "pytest knobs xfuse rc=$?"; tail -n 2 gpurun_out/r2am_pytest_knobs.log
for rep in 1; do for xf in 0 1 0 1; do
MOGP_XFUSE=$xf timeout 300 python bench.py --steps 300 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('xfuse=$xf rep$rep value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"
done; done
for xf in 0 1; do MOGP_XFUSE=$xf DIAG_CFGS=cfg1,cfg4,cfg3 ROWP_COMBOS="1:4096:1:1:0" timeout 300 python tools/gpu_diag.py rowp 2>&1 | grep "step" | sed "s/^/xfuse=$xf /"; done
