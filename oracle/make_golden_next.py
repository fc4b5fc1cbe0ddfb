#!/usr/bin/env python
"""Pin oracle/next_kernels.py against the LIVE reference and write tests/golden/next_*.npz.  TEST INFRASTRUCTURE;
runs only in the build container (needs /root/reference).  Usage: python oracle/make_golden_next.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.make_golden import import_reference, GOLDEN      # noqa: E402
from oracle import next_kernels as nk                        # noqa: E402


def main():
    mogptk = import_reference()
    gpr = mogptk.gpr
    torch.manual_seed(3)
    rng = np.random.default_rng(3)
    for kind, C, Q, Rq, D, ns in [("CSM", 3, 2, 2, 1, [17, 9, 12]), ("CSM", 2, 3, 1, 2, [11, 14]),
                                  ("SMLMC", 3, 2, 2, 1, [13, 8, 10]), ("SMLMC", 2, 2, 3, 2, [9, 12]),
                                  ("UMOSM", 3, 2, 1, 1, [12, 9, 11]), ("UMOSM", 2, 2, 1, 2, [10, 8]),
                                  ("MOHSM", 3, 2, 1, 1, [11, 10, 9]), ("MOHSM", 2, 1, 1, 2, [9, 8])]:
        xs = [torch.tensor(np.sort(rng.uniform(0, 4, (n, D)), axis=0)) for n in ns]
        if kind == "CSM":
            kernel = gpr.MixtureKernel(gpr.CrossSpectralKernel(output_dims=C, input_dims=D, Rq=Rq), Q)
            p = {"amplitude": torch.rand(Q, C, Rq, dtype=torch.float64) + 0.2,
                 "mean": torch.rand(Q, D, dtype=torch.float64) + 0.05,
                 "variance": torch.rand(Q, D, dtype=torch.float64) + 0.1,
                 "shift": 0.3 * torch.randn(Q, C, Rq, dtype=torch.float64)}
            for q in range(Q):
                kernel[q].amplitude.assign(p["amplitude"][q])
                kernel[q].mean.assign(p["mean"][q])
                kernel[q].variance.assign(p["variance"][q])
                kernel[q].shift.assign(p["shift"][q])
            # what the reference actually holds (constrained values after assign)
            p = {"amplitude": torch.stack([kernel[q].amplitude().detach() for q in range(Q)]),
                 "mean": torch.stack([kernel[q].mean().detach() for q in range(Q)]),
                 "variance": torch.stack([kernel[q].variance().detach() for q in range(Q)]),
                 "shift": torch.stack([kernel[q].shift().detach() for q in range(Q)])}
        elif kind in ("UMOSM", "MOHSM"):
            cls = gpr.UncoupledMultiOutputSpectralKernel if kind == "UMOSM" else gpr.MultiOutputHarmonizableSpectralKernel
            kernel = gpr.MixtureKernel(cls(output_dims=C, input_dims=D), Q)
            names = nk.PARAM_NAMES[kind]
            for q in range(Q):
                if kind == "UMOSM":
                    kernel[q].weight.assign((torch.rand(C, C, dtype=torch.float64) + 0.3).tril())
                else:
                    kernel[q].weight.assign(torch.rand(C, dtype=torch.float64) + 0.3)
                    kernel[q].lengthscale.assign(torch.rand(C, dtype=torch.float64) * 0.5 + 0.2)
                    kernel[q].center.assign(torch.rand(D, dtype=torch.float64) * 2.0 + 1.0)
                kernel[q].mean.assign(torch.rand(C, D, dtype=torch.float64) + 0.05)
                kernel[q].variance.assign(torch.rand(C, D, dtype=torch.float64) + 0.1)
                kernel[q].delay.assign(0.2 * torch.randn(C, D, dtype=torch.float64))
                kernel[q].phase.assign(0.4 * torch.randn(C, dtype=torch.float64))
            p = {k: torch.stack([getattr(kernel[q], k)().detach() for q in range(Q)]) for k in names}
        else:
            kernel = gpr.LinearModelOfCoregionalizationKernel([gpr.SpectralKernel(D) for _ in range(Q)], output_dims=C,
                                                              input_dims=D, Q=Q, Rq=Rq)
            kernel.weight.assign(torch.rand(C, Q, Rq, dtype=torch.float64) + 0.1)
            for q in range(Q):
                kernel[q].magnitude.assign(torch.rand(1, dtype=torch.float64) + 0.2)
                kernel[q].mean.assign(torch.rand(D, dtype=torch.float64) + 0.05)
                kernel[q].variance.assign(torch.rand(D, dtype=torch.float64) * 0.2 + 0.02)
            p = {"weight": kernel.weight().detach(),
                 "magnitude": torch.stack([kernel[q].magnitude().detach().reshape(()) for q in range(Q)]),
                 "mean": torch.stack([kernel[q].mean().detach() for q in range(Q)]),
                 "variance": torch.stack([kernel[q].variance().detach() for q in range(Q)])}
        X = torch.cat([torch.cat([torch.full((x.shape[0], 1), float(c), dtype=torch.float64), x], dim=1)
                       for c, x in enumerate(xs)])
        with torch.no_grad():
            Kref = kernel.K(X).detach()
            kd_ref = kernel.K_diag(X).detach()
        # restatement == reference, block by block
        off = np.cumsum([0] + ns)
        worst = 0.0
        for i in range(C):
            for j in range(C):
                blk = nk.KSUB[kind](i, j, xs[i], xs[j], p)
                worst = max(worst, float((blk - Kref[off[i]:off[i + 1], off[j]:off[j + 1]]).abs().max()))
        assert worst <= 1e-13 * float(Kref.abs().max()), (kind, worst)
        kd_err = float((kd_ref - Kref.diagonal()).abs().max())
        # reference quirk: SpectralKernel.K sums exp*cos over the input dimensions (singleoutput.py:556) while its K_diag
        # returns the bare magnitude (:558-561), so K_diag != diag(K) for D > 1 (the SM kernel shares it)
        assert kd_err < 1e-13 * float(Kref.abs().max()) or (kind == "SMLMC" and D > 1), (kind, D, kd_err)
        name = "next_%s_c%dq%dr%dd%d" % (kind.lower(), C, Q, Rq, D)
        out = {"kind": kind, "C": C, "Q": Q, "Rq": Rq, "D": D, "X": X.numpy(), "K": Kref.numpy(), "K_diag": kd_ref.numpy()}
        out.update({"p_" + k: v.numpy() for k, v in p.items()})
        # the exact-GP step of the reference on this kernel: LML, gradients w.r.t. the constrained values, predictions
        orc = nk.register()
        y = torch.tensor(rng.standard_normal((X.shape[0], 1)))
        sigma = torch.tensor(0.3 + 0.4 * rng.uniform(size=C))
        jitter = 1e-8
        model = gpr.Exact(kernel, X.numpy(), y.numpy(), variance=(sigma ** 2).tolist(), jitter=jitter)
        sigma = model.likelihood.scale().detach().reshape(-1).clone()      # what the reference holds after its transform
        lml_ref = float(model.log_marginal_likelihood().detach())
        lml_orc = float(orc.lml(kind, p, sigma, X, y, jitter))
        assert abs(lml_orc - lml_ref) <= 1e-12 * abs(lml_ref), (kind, lml_orc, lml_ref)
        _, g_orc = orc.loss_and_grad(kind, p, sigma, X, y, jitter)
        # reference gradients w.r.t. the constrained values: autograd through its own kernel objects
        cons = {}
        if kind == "CSM":
            leaves = {k: [getattr(kernel[q], k) for q in range(Q)] for k in ("amplitude", "mean", "variance", "shift")}
        elif kind in ("UMOSM", "MOHSM"):
            leaves = {k: [getattr(kernel[q], k) for q in range(Q)] for k in nk.PARAM_NAMES[kind]}
        else:
            leaves = {"weight": [kernel.weight], "magnitude": [kernel[q].magnitude for q in range(Q)],
                      "mean": [kernel[q].mean for q in range(Q)], "variance": [kernel[q].variance for q in range(Q)]}
        model.zero_grad()
        loss = -model.log_marginal_likelihood()
        flat = [t for ts in leaves.values() for t in ts]
        cvals = [t() for t in flat]                                   # constrained values (graph nodes)
        grads = torch.autograd.grad(-model.log_marginal_likelihood(), flat, allow_unused=True)
        # chain back from raw to constrained: d/d constrained = d/d raw / (d constrained / d raw)
        gi = 0
        for kname, ts in leaves.items():
            parts = []
            for t in ts:
                c = t()
                (dc,) = torch.autograd.grad(c.sum(), t, retain_graph=True)
                g_raw = grads[gi] if grads[gi] is not None else torch.zeros_like(t)
                parts.append((g_raw / dc).detach().reshape(c.shape))
                gi += 1
            ref = parts[0] if (kname == "weight" and kind == "SMLMC") else torch.stack([x.reshape(x.shape) for x in parts])
            if kind == "SMLMC" and kname == "magnitude":
                ref = ref.reshape(-1)
            got = g_orc[kname]
            scale = max(float(ref.abs().max()), 1e-12)
            assert float((got.reshape(ref.shape) - ref).abs().max()) <= 2e-8 * scale, (kind, kname, got, ref)
            out["gc_" + kname] = got.numpy()
        out["gc_sigma"] = g_orc["sigma"].numpy()
        Xs = torch.cat([torch.cat([torch.full((5, 1), float(c), dtype=torch.float64),
                                   torch.tensor(rng.uniform(0, 4, (5, D)))], dim=1) for c in range(C)])
        mu_ref, var_ref = model.predict_f(Xs.numpy())
        mu_orc, var_orc = orc.predict_f(kind, p, sigma, X, y, Xs, jitter)
        assert float((mu_orc - mu_ref).abs().max()) <= 1e-9 * max(float(mu_ref.abs().max()), 1e-12)
        assert float((var_orc - var_ref).abs().max()) <= 1e-9 * max(float(var_ref.abs().max()), 1e-12)
        out.update({"y": y.numpy().reshape(-1), "sigma": sigma.numpy(), "jitter": jitter, "lml": lml_ref, "Xs": Xs.numpy(),
                    "pred_mu": mu_ref.detach().numpy().reshape(-1), "pred_var": var_ref.detach().numpy().reshape(-1)})
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
        print("%s: restatement == reference (max abs diff %.1e), |K_diag - diag K| = %.1e, wrote %s.npz" % (kind, worst, kd_err, name))


if __name__ == "__main__":
    main()
