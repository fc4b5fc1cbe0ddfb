"""Synthetic workloads for the BASELINE.json configs.

Pure numpy/torch-CPU host code: generates the data (kernel format) and a set of
constrained hyper-parameters.  The parameter draw follows the order of the
reference constructors (mogptk/models/mosm.py:53-55, sm.py:52-54, conv.py:54-56:
``torch.rand`` after ``torch.manual_seed(seed)``) followed by the explicit mean
assignment the survey prescribes, and each value goes through the reference's
assign -> inverse -> forward transform round trip (gpr/parameter.py:59) so the
values equal what the reference model would hold.
"""
import math

import numpy as np
import torch

CONFIGS = {
    # name: (kind, C, n_per_channel, Q)
    "cfg1": ("SM", 1, 512, 3),
    "cfg2": ("MOSM", 4, 512, 5),
    "cfg3": ("MOSM", 8, 1024, 10),
    "cfg4": ("CONV", 4, 1024, 1),
}

_LO = 1e-8   # gpr/config.py positive_minimum


def make_data(C, n_per_channel, seed=0, D=1):
    """x_c = sort(U(0,10)), y_c = two sinusoids + noise; returns X (N,1+D), y (N,)."""
    rng = np.random.default_rng(seed)
    if np.isscalar(n_per_channel):
        n_per_channel = [int(n_per_channel)] * C
    xs, ys = [], []
    for c in range(C):
        n = n_per_channel[c]
        x = np.sort(rng.uniform(0.0, 10.0, n))
        y = np.sin(2 * np.pi * (0.3 + 0.2 * c) * x) + 0.5 * np.sin(2 * np.pi * 1.1 * x + c) \
            + 0.1 * rng.standard_normal(n)
        cols = [np.full(n, float(c)), x]
        for d in range(1, D):
            cols.append(rng.uniform(0.0, 10.0, n))
        xs.append(np.stack(cols, axis=1))
        ys.append(y)
    return np.concatenate(xs, axis=0), np.concatenate(ys, axis=0)


def _roundtrip(v, lower=_LO, beta=0.1):
    """assign(value) then read back: softplus(inverse(value)) with the reference's
    inverse (parameter.py:59) -- not an exact identity."""
    v = v.to(torch.float64)
    raw = (v - lower) + torch.log(-torch.expm1(-beta * v - lower)) / beta
    return lower + torch.nn.functional.softplus(raw, beta=beta, threshold=20.0)


def make_params(kind, C, Q, D=1, seed=0, random_delay_phase=False):
    """Constrained hyper-parameters (dict of fp64 tensors) + noise scale sigma (C,)."""
    torch.manual_seed(seed)
    g = torch.rand
    if kind == "MOSM":
        w = g(C, Q); _ = g(C, Q, D); v = g(C, Q, D)
        mu = g(C, Q, 1).expand(C, Q, D).clone() * 2.0 + 0.05 if D == 1 else g(C, Q, D) * 2.0 + 0.05
        p = {"weight": _roundtrip(w), "mean": _roundtrip(mu), "variance": _roundtrip(v),
             "delay": torch.zeros(C, Q, D, dtype=torch.float64),
             "phase": torch.zeros(C, Q, dtype=torch.float64)}
        if random_delay_phase:
            p["delay"] = 0.3 * torch.randn(C, Q, D, dtype=torch.float64)
            p["phase"] = 0.5 * torch.randn(C, Q, dtype=torch.float64)
    elif kind == "SM":
        mags, means, vars_ = [], [], []
        for _c in range(C):
            mags.append(g(Q)); _ = g(Q, D); vars_.append(g(Q, D))
        for _c in range(C):
            means.append(g(Q, D) * 2.0 + 0.05)
        p = {"magnitude": _roundtrip(torch.stack(mags)), "mean": _roundtrip(torch.stack(means)),
             "variance": _roundtrip(torch.stack(vars_))}
    elif kind == "CONV":
        ws, vs, bs = [], [], []
        for _q in range(Q):
            ws.append(g(C)); vs.append(g(C, D)); bs.append(g(D))
        p = {"weight": _roundtrip(torch.stack(ws)),
             "variance": _roundtrip(torch.stack(vs), lower=0.0),
             "base_variance": _roundtrip(torch.stack(bs))}
    else:
        raise ValueError("unknown kernel kind %r" % (kind,))
    sigma = _roundtrip(torch.ones(C))
    return p, sigma


def make_config(name, seed=0):
    kind, C, n, Q = CONFIGS[name]
    X, y = make_data(C, n, seed)
    p, sigma = make_params(kind, C, Q, 1, seed)
    return kind, p, sigma, X, y


def n_params(kind, C, Q, D=1):
    if kind == "MOSM":
        return C * Q * (2 + 3 * D) + C
    if kind == "SM":
        return C * Q * (1 + 2 * D) + C
    return Q * (C + C * D + D) + C


def flops_per_iteration(N):
    """potrf N^3/3 + inverse 2N^3/3 (+ O(N^2))."""
    return float(N) ** 3
