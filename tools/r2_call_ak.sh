#!/bin/bash
# A/B on ONE box: the tree of commit 00b217a (before early loss / z along the pipeline) against the current tree
mkdir -p gpurun_out
for rep in 1 2; do
(cd ab_old && timeout 300 python bench.py --steps 300 --no-extras --no-cpu-baseline 2>/dev/null) | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('OLD rep$rep value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"
timeout 300 python bench.py --steps 300 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NEW rep$rep value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))"
done
