#!/usr/bin/env python
"""Pin the oracle against the LIVE reference and write the golden fixtures.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference,
which does not exist on the GPU box).  For every case it

  1. builds the reference model (mogptk.gpr kernels + gpr.Exact) on CPU/fp64,
  2. reads back the constrained parameter values the reference holds,
  3. checks oracle K / K_diag / LML / raw-space gradients / predict_f against the
     reference (asserts; tolerances below),
  4. writes tests/golden/<case>.npz with inputs and reference outputs.

Usage:  python oracle/make_golden.py [--cases a,b,...] [--big]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference():
    """the reference imports matplotlib/IPython at module import -- see oracle/ref_loader.py."""
    from oracle import ref_loader
    return ref_loader.import_reference("cpu")


from oracle import mogp_oracle as orc            # noqa: E402
from mogptk_b200 import synth                     # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_reference_model(mogptk, kind, C, Q, D, X, y, pvals, sigma, jitter, data_var=None):
    gpr = mogptk.gpr
    if kind == "MOSM":
        kernel = gpr.MultiOutputSpectralMixtureKernel(Q=Q, output_dims=C, input_dims=D)
        for k in ("weight", "mean", "variance", "delay", "phase"):
            getattr(kernel, k).assign(pvals[k])
        plist = {k: getattr(kernel, k) for k in ("weight", "mean", "variance", "delay", "phase")}
    elif kind == "SM":
        kernel = gpr.IndependentMultiOutputKernel(
            [gpr.SpectralMixtureKernel(Q=Q, input_dims=D) for _ in range(C)], output_dims=C)
        for c in range(C):
            kernel[c].magnitude.assign(pvals["magnitude"][c])
            kernel[c].mean.assign(pvals["mean"][c])
            kernel[c].variance.assign(pvals["variance"][c])
        plist = None
    elif kind == "CONV":
        kernel = gpr.MixtureKernel(gpr.GaussianConvolutionProcessKernel(output_dims=C, input_dims=D), Q)
        for q in range(Q):
            kernel[q].weight.assign(pvals["weight"][q])
            kernel[q].variance.assign(pvals["variance"][q])
            kernel[q].base_variance.assign(pvals["base_variance"][q])
        plist = None
    model = gpr.Exact(kernel, X, y, variance=(sigma ** 2).tolist(), data_variance=data_var, jitter=jitter)
    return model


def read_back(kind, model, C, Q):
    """Constrained values + handles to the raw Parameter objects, oracle layout."""
    k = model.kernel
    if kind == "MOSM":
        objs = {n: [getattr(k, n)] for n in ("weight", "mean", "variance", "delay", "phase")}
        stack = lambda lst: lst[0]
    elif kind == "SM":
        objs = {n: [getattr(k[c], n) for c in range(C)] for n in ("magnitude", "mean", "variance")}
        stack = lambda lst: torch.stack(lst)
    else:
        objs = {n: [getattr(k[q], n) for q in range(Q)] for n in ("weight", "variance", "base_variance")}
        stack = lambda lst: torch.stack(lst)
    cons = {n: stack([o().detach().clone() for o in lst]) for n, lst in objs.items()}
    objs["sigma"] = [model.likelihood.scale]
    cons_sigma = model.likelihood.scale().detach().clone()
    return cons, cons_sigma, objs, stack


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / den)


def run_case(mogptk, name, kind, C, ns, Q, D=1, seed=0, rdp=True, sigma=None, jitter=1e-8,
             shuffle=False, with_data_var=False, n_pred=48, full_K_max=96, grads=True, pred=True):
    t0 = time.time()
    X, y = synth.make_data(C, ns, seed, D)
    pvals, sig = synth.make_params(kind, C, Q, D, seed, random_delay_phase=rdp)
    if sigma is not None:
        sig = torch.tensor(sigma, dtype=torch.float64)
    if shuffle:
        perm = np.random.default_rng(seed + 100).permutation(X.shape[0])
        X, y = X[perm], y[perm]
    data_var = None
    if with_data_var:
        data_var = 0.05 + 0.1 * np.random.default_rng(seed + 7).uniform(size=X.shape[0])
    N = X.shape[0]
    model = build_reference_model(mogptk, kind, C, Q, D, X, y, pvals, sig, jitter, data_var)
    cons, cons_sigma, objs, stack = read_back(kind, model, C, Q)
    out = dict(kind=kind, C=C, Q=Q, D=D, X=X, y=y, jitter=jitter, sigma=cons_sigma.numpy())
    if data_var is not None:
        out["data_var"] = data_var
    for n, v in cons.items():
        out["p_" + n] = v.numpy()
    Xt = torch.tensor(X, dtype=torch.float64)

    # ---- K, K_diag
    with torch.no_grad():
        Kref = model.kernel.K(Xt)
        Kor = orc.K(kind, cons, Xt)
        e = relerr(Kor, Kref)
        assert e < 1e-13, (name, "K", e)
        kd_ref = model.kernel.K_diag(Xt)
        kd_or = orc.K_diag(kind, cons, Xt)
        assert relerr(kd_or, kd_ref) < 1e-14, (name, "K_diag")
    Kref = Kref.numpy()
    if N <= full_K_max:
        out["K_full"] = Kref
    else:
        rng = np.random.default_rng(seed + 1)
        S = 4096
        idx = np.stack([rng.integers(0, N, S), rng.integers(0, N, S)], axis=1)
        idx[:256, 1] = idx[:256, 0]                       # include diagonal entries
        idx[256:512, 1] = np.clip(idx[256:512, 0] + rng.integers(-3, 4, 256), 0, N - 1)   # near-diagonal
        out["K_idx"] = idx
        out["K_val"] = Kref[idx[:, 0], idx[:, 1]]
        out["K_rowsum0"] = Kref[:, 0].copy()
        out["K_fro"] = float(np.sqrt((Kref ** 2).sum()))
    out["K_diag"] = kd_ref.numpy()

    # ---- LML
    with torch.no_grad():
        lml_ref = float(model.log_marginal_likelihood())
    lml_or = float(orc.lml(kind, cons, cons_sigma, Xt, y, jitter, data_var))
    e = abs(lml_or - lml_ref) / abs(lml_ref)
    assert e < 1e-12, (name, "lml", lml_or, lml_ref)
    out["lml"] = lml_ref

    # ---- loss + gradients (raw space from the reference; constrained from the oracle)
    if grads:
        loss_ref = model.loss()
        out["loss"] = float(loss_ref.detach())
        loss_or, g_or = orc.loss_and_grad(kind, cons, cons_sigma, Xt, y, jitter, data_var)
        assert abs(float(loss_or) - float(loss_ref)) / abs(float(loss_ref)) < 1e-12
        for n, lst in objs.items():
            raw = stack(lst) if n != "sigma" else lst[0]
            raw = raw.detach()
            zg = lambda o: o.grad if o.grad is not None else torch.zeros_like(o)
            graw = (stack([zg(o) for o in lst]) if n != "sigma" else zg(lst[0])).detach()
            out["r_" + n] = raw.numpy()
            out["gr_" + n] = graw.numpy()
            out["gc_" + n] = g_or[n].numpy()
            # chain the oracle's constrained gradient to raw space and compare with the reference
            lo = lst[0].lower
            up = lst[0].upper
            r = raw.clone().requires_grad_(True)
            if lo is not None and up is None:
                assert lo.ndim == 0
                c = orc.softplus_forward(r, float(lo))
            elif lo is None and up is None:
                c = r
            else:
                raise AssertionError("unexpected bounds in golden case")
            (c * g_or[n]).sum().backward()
            scale = max(float(graw.abs().max()), 1e-12)
            e = float((r.grad - graw).abs().max()) / scale
            assert e < 2e-8, (name, "grad", n, e)
            out["lower_" + n] = np.asarray(0.0 if lo is None else float(lo.reshape(-1)[0]))
            out["has_lower_" + n] = np.asarray(lo is not None)

    # ---- prediction
    if pred:
        rng = np.random.default_rng(seed + 2)
        D_ = D
        cs = rng.integers(0, C, n_pred).astype(np.float64)
        xs = rng.uniform(-0.5, 10.5, (n_pred, D_))
        Xs = np.concatenate([cs[:, None], xs], axis=1)
        mu_ref, var_ref = model.predict_f(torch.tensor(Xs))
        mu_or, var_or = orc.predict_f(kind, cons, cons_sigma, Xt, y, Xs, jitter, data_var=data_var)
        assert relerr(mu_or, mu_ref) < 1e-9, (name, "pred mu", relerr(mu_or, mu_ref))
        assert relerr(var_or, var_ref) < 1e-9, (name, "pred var", relerr(var_or, var_ref))
        out.update(Xs=Xs, pred_mu=mu_ref.numpy().reshape(-1), pred_var=var_ref.numpy().reshape(-1))
        if N <= 2048:
            _, cov_ref = model.predict_f(torch.tensor(Xs), full=True)
            _, cov_or = orc.predict_f(kind, cons, cons_sigma, Xt, y, Xs, jitter, full=True, data_var=data_var)
            assert relerr(cov_or, cov_ref) < 1e-9
            out["pred_cov"] = cov_ref.numpy()
        with torch.no_grad():
            Kfs_ref = model.kernel.K(Xt, torch.tensor(Xs)).numpy()
        assert relerr(orc.K(kind, cons, Xt, Xs), Kfs_ref) < 1e-13
        out["Kfs_rows"] = Kfs_ref[:: max(1, N // 64)].copy()
        out["Kfs_row_stride"] = max(1, N // 64)

    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("%-18s kind=%-4s N=%-5d lml=%.9f  ok  (%.1fs)" % (name, kind, N, out["lml"], time.time() - t0), flush=True)


CASES = {
    # name: kwargs
    "mosm_small":   dict(kind="MOSM", C=3, ns=[17, 9, 13], Q=2, sigma=[0.3, 0.7, 0.5]),
    "mosm_small_d2": dict(kind="MOSM", C=2, ns=[21, 30], Q=3, D=2, sigma=[0.4, 0.6], seed=3),
    "mosm_shuffled": dict(kind="MOSM", C=3, ns=[40, 25, 70], Q=2, sigma=[0.2, 0.3, 0.25], shuffle=True, seed=5),
    "mosm_datavar": dict(kind="MOSM", C=2, ns=[33, 47], Q=2, sigma=[0.3, 0.2], with_data_var=True, seed=6),
    "mosm_c1":      dict(kind="MOSM", C=1, ns=[90], Q=3, sigma=[0.1], seed=8),
    "mosm_mid":     dict(kind="MOSM", C=4, ns=[100, 77, 130, 64], Q=3, sigma=[0.15, 0.2, 0.1, 0.3], seed=2),
    "sm_small":     dict(kind="SM", C=2, ns=[31, 18], Q=3, sigma=[0.3, 0.5], seed=1),
    "sm_small_d2":  dict(kind="SM", C=2, ns=[25, 25], Q=2, D=2, sigma=[0.3, 0.5], seed=4),
    "conv_small":   dict(kind="CONV", C=3, ns=[20, 31, 12], Q=2, sigma=[0.3, 0.4, 0.5], seed=1),
    "conv_small_d2": dict(kind="CONV", C=2, ns=[26, 22], Q=2, D=2, sigma=[0.3, 0.4], seed=9),
    "cfg1":         dict(kind="SM", C=1, ns=512, Q=3, rdp=False),
    "cfg2":         dict(kind="MOSM", C=4, ns=512, Q=5, rdp=False),
    "cfg2_rdp":     dict(kind="MOSM", C=4, ns=512, Q=5, rdp=True, seed=1, sigma=[0.2, 0.3, 0.25, 0.15]),
    "cfg4":         dict(kind="CONV", C=4, ns=1024, Q=1, rdp=False),
}
BIG = {
    "cfg3":         dict(kind="MOSM", C=8, ns=1024, Q=10, rdp=False, grads=True, pred=False),
}


def add_prediction(mogptk, name, n_pred=48):
    """Append predict_f outputs of the live reference to an existing fixture (forward only: used for cfg3, whose
    gradients take minutes and tens of GB through autograd but whose posterior needs one K build + Cholesky)."""
    t0 = time.time()
    path = os.path.join(GOLDEN, name + ".npz")
    z = np.load(path, allow_pickle=False)
    out = {k: z[k] for k in z.files}
    kind, C, Q, D = str(out["kind"]), int(out["C"]), int(out["Q"]), int(out["D"])
    X, y, jitter = out["X"], out["y"], float(out["jitter"])
    cons = {k[2:]: torch.tensor(v, dtype=torch.float64) for k, v in out.items() if k.startswith("p_")}
    sig = torch.tensor(out["sigma"], dtype=torch.float64)
    model = build_reference_model(mogptk, kind, C, Q, D, X, y, cons, sig, jitter, out.get("data_var"))
    # the fixture was made from the raw values: restore them exactly (assign() is not an exact round trip)
    _, _, objs, stack = read_back(kind, model, C, Q)
    for n, lst in objs.items():
        raw = torch.tensor(out["r_" + n], dtype=torch.float64)
        for i, prm in enumerate(lst):
            prm.data = (raw if len(lst) == 1 else raw[i]).clone().reshape(prm.shape)
    with torch.no_grad():
        lml = float(model.log_marginal_likelihood())
    assert abs(lml - float(out["lml"])) <= 1e-12 * abs(lml), (lml, float(out["lml"]))
    rng = np.random.default_rng(2)
    cs = rng.integers(0, C, n_pred).astype(np.float64)
    xs = rng.uniform(-0.5, 10.5, (n_pred, D))
    Xs = np.concatenate([cs[:, None], xs], axis=1)
    with torch.no_grad():
        mu_ref, var_ref = model.predict_f(torch.tensor(Xs))
    cons_rb, sig_rb, _, _ = read_back(kind, model, C, Q)
    mu_or, var_or = orc.predict_f(kind, cons_rb, sig_rb, torch.tensor(X), y, Xs, jitter, data_var=out.get("data_var"))
    assert relerr(mu_or, mu_ref) < 1e-9 and relerr(var_or, var_ref) < 1e-8, (relerr(mu_or, mu_ref), relerr(var_or, var_ref))
    out.update(Xs=Xs, pred_mu=mu_ref.numpy().reshape(-1), pred_var=var_ref.numpy().reshape(-1))
    np.savez_compressed(path, **out)
    print("%-18s predictions added (%d points, %.1fs)" % (name, n_pred, time.time() - t0), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="")
    ap.add_argument("--add-pred", default="", help="append reference predictions to these existing fixtures")
    ap.add_argument("--big", action="store_true", help="also cfg3 (N=8192; ~minutes and tens of GB of RAM)")
    args = ap.parse_args()
    mogptk = import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    if args.add_pred:
        for name in args.add_pred.split(","):
            add_prediction(mogptk, name)
        return
    cases = dict(CASES)
    if args.big:
        cases.update(BIG)
    sel = [c for c in args.cases.split(",") if c] or list(cases)
    for name in sel:
        kw = dict({**CASES, **BIG}[name])
        run_case(mogptk, name, **kw)


if __name__ == "__main__":
    main()
