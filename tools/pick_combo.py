#!/usr/bin/env python
"""Pick the fastest (panel variant, inverse pipeline) combination whose numerics passed, from the logs of
`tools/gpu_diag.py exp` (gpurun_out/exp_v*.log).  Prints "variant pipe"; "0 0" when nothing else qualifies."""
import glob
import re
import sys

best, best_t = (0, 0), None
ok, bad, t_cfg2 = set(), set(), {}
for f in glob.glob(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/exp_v*.log"):
    for line in open(f, errors="replace"):
        m = re.match(r"\[v(\d) pipe(\d)\] (.*)", line)
        if not m:
            continue
        key, rest = (int(m.group(1)), int(m.group(2))), m.group(3)
        if rest.startswith("VERDICT"):
            (ok if rest.startswith("VERDICT OK") else bad).add(key)
        if rest.startswith("step"):
            if "STEP_FAIL" in rest:
                bad.add(key)
            mm = re.match(r"step cfg2\s+N=\d+: ([\d.]+) ms", rest)
            if mm:
                t_cfg2[key] = float(mm.group(1))
for key, t in sorted(t_cfg2.items()):
    if key in ok and key not in bad and (best_t is None or t < best_t):
        best, best_t = key, t
print("%d %d" % best)
