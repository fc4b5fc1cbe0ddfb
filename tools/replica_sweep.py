#!/usr/bin/env python
"""Aggregate it/s of R concurrent replicas of one BASELINE config on ONE GPU (R from the command line)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
for R in [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "2,3,4,6,8").split(",")]:
    r = bench.concurrent_replicas(cfg, 0, R, 128)
    print(json.dumps({"cfg": cfg, "replicas": R, "aggregate_it_s": round(r["value"], 1), "single": round(r["single_replica_value"], 1),
                      "ratio": round(r["value"] / r["single_replica_value"], 3)}), flush=True)
