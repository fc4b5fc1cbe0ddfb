#!/usr/bin/env python
"""Where a Python-level training iteration at cfg2 spends its time (host side): loss() pieces, optimiser step, read-back."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                   # noqa: E402
import mogptk_b200 as mb                        # noqa: E402

mb.gpr.use_gpu(0)
model, _, _ = bench.mirror_model("cfg2", 0)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
optf = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
for _ in range(5):
    model.loss(); opt.step()
torch.cuda.synchronize()


def timeit(fn, n=300):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("loss() alone                          %.1f us" % timeit(lambda: model.loss()))
print("_fast_table() alone                   %.1f us" % timeit(lambda: model._fast_table()))
print("opt.step() foreach (async)            %.1f us" % timeit(lambda: opt.step()))
print("opt.step() fused (async)              %.1f us" % timeit(lambda: optf.step()))
print("loss() + foreach step + float         %.1f us" % timeit(lambda: (float(model.loss()), opt.step())))
print("loss() + fused step + float           %.1f us" % timeit(lambda: (float(model.loss()), optf.step())))
eng = model._eng()
rows = model._rows
from mogptk_b200.engine import pack_params
kind, p, _ = mb.gpr.kernel_spec(model.kernel)
packed = pack_params(kind, {k: v.detach() for k, v in p.items()}, eng.device)
sig = model._sigma().detach().contiguous()
print("engine.lml_grad_prepared (async)      %.1f us" % timeit(lambda: eng.lml_grad_prepared(rows, packed, sig, 1e-8, True, check=False)))
print("fit_adam per iteration (64 per sync)  %.1f us" % (timeit(lambda: mb.fit_adam(model, 64, lr=1e-3, sync_every=64), n=5) / 64))
