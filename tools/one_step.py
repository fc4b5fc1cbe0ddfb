#!/usr/bin/env python
"""A few exact-GP steps of one BASELINE config on cuda:0 (for ncu captures)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mogptk_b200 import synth                      # noqa: E402
from mogptk_b200.engine import Engine, pack_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
kind, p, sigma, X, y = synth.make_config(a.config, 0)
eng = Engine(device=0, max_n=X.shape[0])
rows = eng.prepare(kind, p, X, y)
packed = pack_params(kind, p, eng.device)
sig = sigma.to(eng.device)
for _ in range(a.steps):
    out = eng.lml_grad_prepared(rows, packed, sig, 1e-8, True, check=False)
torch.cuda.synchronize()
print("lml", float(out[0]), "info", float(out[1]))
