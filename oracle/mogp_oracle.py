"""CPU oracle for the exact multi-output GP hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain torch-CPU / fp64 restatement of the reference algorithm
(GAMES-UChile/mogptk v0.5.1).  It exists so that tests can check the CUDA path
against it; it is NOT part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product (``mogptk_b200``) never imports it
and has no CPU fallback.

Parity pin: the reference holds no golden vectors for this path, so
the pin is the live reference: ``oracle/make_golden.py`` imports
``/root/reference`` in the build container, checks every function below against
it, and commits fixtures under ``tests/golden/``.

Like the reference the gradient is obtained with autograd through the
materialised (Q, n, m) temporaries, so timing this module reproduces the
reference's cost structure (the ``port`` CPU baseline).

Each function cites the reference file:line it restates.

Parameter layouts (constrained values, fp64):
  MOSM  weight (C,Q)  mean (C,Q,D)  variance (C,Q,D)  delay (C,Q,D)  phase (C,Q)
  SM    magnitude (C,Q)  mean (C,Q,D)  variance (C,Q,D)      (one SM kernel per channel)
  CONV  weight (Q,C)  variance (Q,C,D)  base_variance (Q,D)  (mixture of Q CONV kernels)
X is in "kernel format": column 0 = channel id (float), columns 1.. = inputs.
"""
import math

import numpy as np
import torch

DT = torch.float64
PI = math.pi

KINDS = ("MOSM", "SM", "CONV")
PARAM_NAMES = {
    "MOSM": ("weight", "mean", "variance", "delay", "phase"),
    "SM": ("magnitude", "mean", "variance"),
    "CONV": ("weight", "variance", "base_variance"),
}


def t64(a):
    if isinstance(a, torch.Tensor):
        return a.to(DT)
    return torch.tensor(np.asarray(a), dtype=DT)


# --------------------------------------------------------------------------------------
# constrained-parameter transforms  (mogptk/gpr/parameter.py:30-96)
# --------------------------------------------------------------------------------------

def softplus_forward(raw, lower, beta=0.1, threshold=20.0):
    """lower + softplus_beta(raw)  (parameter.py:48-49)."""
    return lower + torch.nn.functional.softplus(raw, beta=beta, threshold=threshold)


def softplus_inverse(y, lower, beta=0.1):
    """The reference's (not exactly inverse) inverse  (parameter.py:59)."""
    return (y - lower) + torch.log(-torch.expm1(-beta * y - lower)) / beta


def sigmoid_forward(raw, lower, upper):
    """lower + (upper-lower)*sigmoid(raw)  (parameter.py:77-78)."""
    return lower + (upper - lower) * torch.sigmoid(raw)


def sigmoid_inverse(y, lower, upper):
    """logit of the rescaled value  (parameter.py:84,96)."""
    s = (y - lower) / (upper - lower)
    return torch.log(s) - torch.log(1 - s)


# --------------------------------------------------------------------------------------
# per channel-pair sub-Grams
# --------------------------------------------------------------------------------------

def _tau(x1, x2):
    """Signed difference (n,m,D)  (mogptk/gpr/kernel.py:172-177)."""
    return x1.unsqueeze(1) - x2.unsqueeze(0)


def mosm_ksub(i, j, x1, x2, p):
    """MOSM block K_ij(x1,x2)  (mogptk/gpr/multioutput.py:178-204)."""
    D = x1.shape[1]
    twopi_pow = (2.0 * PI) ** (D / 2.0)                        # multioutput.py:176
    tau = _tau(x1, x2)                                          # n,m,D
    w, mu, var = p["weight"], p["mean"], p["variance"]
    if i == j:                                                  # multioutput.py:182-187
        alpha = w[i] ** 2 * twopi_pow * var[i].prod(dim=1).sqrt()
        ex = torch.exp(-0.5 * torch.einsum("nmd,qd->qnm", tau ** 2, var[i]))
        cs = torch.cos(2.0 * PI * torch.einsum("nmd,qd->qnm", tau, mu[i]))
        return (alpha[:, None, None] * ex * cs).sum(dim=0)
    iv = 1.0 / (var[i] + var[j])                                # multioutput.py:189
    dm = mu[i] - mu[j]
    mag = w[i] * w[j] * torch.exp(-PI ** 2 * (dm * iv * dm).sum(dim=1))      # :192
    m = iv * (var[i] * mu[j] + var[j] * mu[i])                  # :194
    v = 2.0 * var[i] * iv * var[j]                              # :195
    th = p["delay"][i] - p["delay"][j]                          # :196
    ph = p["phase"][i] - p["phase"][j]                          # :197
    alpha = mag * twopi_pow * v.prod(dim=1).sqrt()              # :199
    td = tau[None] + th[:, None, None, :]                       # Q,n,m,D  :200
    ex = torch.exp(-0.5 * torch.einsum("qnmd,qd->qnm", td ** 2, v))
    cs = torch.cos(2.0 * PI * (torch.einsum("qnmd,qd->qnm", td, m) + ph[:, None, None]))
    return (alpha[:, None, None] * ex * cs).sum(dim=0)


def mosm_ksub_diag(i, n, p):
    """multioutput.py:206-210."""
    D = p["variance"].shape[2]
    alpha = p["weight"][i] ** 2 * (2.0 * PI) ** (D / 2.0) * p["variance"][i].prod(dim=1).sqrt()
    return alpha.sum().repeat(n)


def sm_ksub(i, j, x1, x2, p):
    """Independent SM kernels (multioutput.py:26-34 -> singleoutput.py:594-600)."""
    if i != j:
        return torch.zeros(x1.shape[0], x2.shape[0], dtype=DT)
    tau = _tau(x1, x2)[None]                                    # 1,n,m,D
    ex = -2.0 * PI ** 2 * tau ** 2 * p["variance"][i][:, None, None, :]
    cs = 2.0 * PI * tau * p["mean"][i][:, None, None, :]
    return torch.einsum("q,qnmd,qnmd->nm", p["magnitude"][i], torch.exp(ex), torch.cos(cs))


def sm_ksub_diag(i, n, p):
    """singleoutput.py:602-605 (sum of magnitudes, whatever D is)."""
    return p["magnitude"][i].sum().repeat(n)


def conv_ksub(i, j, x1, x2, p):
    """Sum over the Q mixture members (kernel.py:242-243) of the CONV block
    (multioutput.py:531-547).  The i==j Gram branch (2*var_i+base) is the same
    formula as the general one."""
    tau2 = _tau(x1, x2) ** 2                                    # n,m,D
    out = torch.zeros(x1.shape[0], x2.shape[0], dtype=DT)
    Q = p["weight"].shape[0]
    for q in range(Q):
        V = p["variance"][q, i] + p["variance"][q, j] + p["base_variance"][q]      # D
        mag = p["weight"][q, i] * p["weight"][q, j] * torch.sqrt(p["base_variance"][q].prod() / V.prod())
        out = out + mag * torch.exp(-0.5 * torch.tensordot(tau2, 1.0 / V, dims=1))
    return out


def conv_ksub_diag(i, n, p):
    """multioutput.py:549-553 summed over the mixture (kernel.py:245-246)."""
    Q = p["weight"].shape[0]
    tot = torch.zeros((), dtype=DT)
    for q in range(Q):
        V = 2.0 * p["variance"][q, i] + p["base_variance"][q]
        tot = tot + p["weight"][q, i] ** 2 * torch.sqrt(p["base_variance"][q].prod() / V.prod())
    return tot.repeat(n)


_KSUB = {"MOSM": mosm_ksub, "SM": sm_ksub, "CONV": conv_ksub}
_KSUB_DIAG = {"MOSM": mosm_ksub_diag, "SM": sm_ksub_diag, "CONV": conv_ksub_diag}


_N_CHANNELS = {"CONV": lambda p: p["weight"].shape[1]}
_DIAG_NEEDS_X = set()      # kinds whose Ksub_diag depends on the inputs (non-stationary: MOHSM): ksub_diag(i, x, p)


def n_channels(kind, p):
    if kind in _N_CHANNELS:
        return _N_CHANNELS[kind](p)
    return p[PARAM_NAMES[kind][0]].shape[0]


def register_kind(kind, param_names, ksub, ksub_diag, n_channels_fn=None, diag_needs_x=False):
    """Let another test-infrastructure module (oracle/next_kernels.py) reuse the block assembly / LML / prediction
    restatements above for a further kernel family."""
    PARAM_NAMES[kind] = tuple(param_names)
    _KSUB[kind] = ksub
    _KSUB_DIAG[kind] = ksub_diag
    if diag_needs_x:
        _DIAG_NEEDS_X.add(kind)
    if n_channels_fn is not None:
        _N_CHANNELS[kind] = n_channels_fn


# --------------------------------------------------------------------------------------
# block assembly  (mogptk/gpr/kernel.py:446-495)
# --------------------------------------------------------------------------------------

def _split(X, C):
    c = X[:, 0].long()
    rows = [torch.nonzero(c == i, as_tuple=False)[:, 0] for i in range(C)]
    xs = [X[r, 1:] for r in rows]
    return rows, xs


def K(kind, p, X1, X2=None):
    """Full kernel matrix from per-pair blocks, any row order  (kernel.py:446-481)."""
    X1 = t64(X1)
    C = n_channels(kind, p)
    ksub = _KSUB[kind]
    r1, x1 = _split(X1, C)
    if X2 is None:
        # lower blocks + transposed copies (kernel.py:458-467); built by block
        # concatenation per channel-sorted order then un-permuted, which is
        # autograd-friendly and equivalent to the reference's index_put scatter.
        blocks = [[None] * C for _ in range(C)]
        for i in range(C):
            for j in range(i + 1):
                k = ksub(i, j, x1[i], x1[j], p)
                blocks[i][j] = k
                if i != j:
                    blocks[j][i] = k.T
        Ks = torch.cat([torch.cat(row, dim=1) for row in blocks], dim=0)
        perm = torch.cat(r1)
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(perm.shape[0])
        return Ks[inv][:, inv]
    X2 = t64(X2)
    r2, x2 = _split(X2, C)
    blocks = [[ksub(i, j, x1[i], x2[j], p) for j in range(C)] for i in range(C)]   # kernel.py:476-479
    Ks = torch.cat([torch.cat(row, dim=1) for row in blocks], dim=0)
    p1, p2 = torch.cat(r1), torch.cat(r2)
    i1 = torch.empty_like(p1); i1[p1] = torch.arange(p1.shape[0])
    i2 = torch.empty_like(p2); i2[p2] = torch.arange(p2.shape[0])
    return Ks[i1][:, i2]


def K_diag(kind, p, X1):
    """kernel.py:483-495."""
    X1 = t64(X1)
    C = n_channels(kind, p)
    rows, xs = _split(X1, C)
    out = torch.empty(X1.shape[0], dtype=DT)
    for i in range(C):
        out[rows[i]] = _KSUB_DIAG[kind](i, xs[i] if kind in _DIAG_NEEDS_X else rows[i].shape[0], p).detach()
    return out


# --------------------------------------------------------------------------------------
# exact GP: log-marginal likelihood, loss+gradient, prediction (mogptk/gpr/model.py)
# --------------------------------------------------------------------------------------

def _noisy_gram(kind, p, sigma, X, jitter, data_var=None):
    """K + diag(sigma_c(r)^2) [+ data_var] + jitter*mean(diag)*I
    (gpr/model.py:439-443 with :183-186 and :242-244)."""
    Kff = K(kind, p, X)
    chan = X[:, 0].long()
    noise = sigma.reshape(-1)
    noise = noise[chan] if noise.numel() > 1 else noise.expand(X.shape[0])
    Kff = Kff + torch.diag(noise.square())
    if data_var is not None:
        Kff = Kff + torch.diag(t64(data_var))
    Kff = Kff + torch.diag((jitter * Kff.diagonal().mean()).repeat(X.shape[0]))
    return Kff


def lml(kind, p, sigma, X, y, jitter=1e-8, data_var=None):
    """log p(y)  (gpr/model.py:438-453).  y already has any mean subtracted."""
    X, y = t64(X), t64(y).reshape(-1, 1)
    Kt = _noisy_gram(kind, p, sigma, X, jitter, data_var)
    L = torch.linalg.cholesky(Kt)                                # :246
    out = -0.5 * X.shape[0] * math.log(2.0 * PI)                 # :436,450
    out = out - L.diagonal().log().sum()                         # :451
    out = out - 0.5 * (y.T @ torch.cholesky_solve(y, L)).squeeze()   # :452
    return out


def loss_and_grad(kind, params, sigma, X, y, jitter=1e-8, data_var=None):
    """loss = -LML and d loss / d (constrained params, sigma) by autograd, as the
    reference's loss() does (gpr/model.py:279-292), minus the transform chain."""
    p = {k: t64(v).clone().requires_grad_(True) for k, v in params.items()}
    s = t64(sigma).clone().requires_grad_(True)
    loss = -lml(kind, p, s, X, y, jitter, data_var)
    loss.backward()
    grads = {k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v)) for k, v in p.items()}
    grads["sigma"] = s.grad.clone()
    return loss.detach(), grads


def predict_f(kind, p, sigma, X, y, Xs, jitter=1e-8, full=False, data_var=None):
    """Posterior mean / variance of f at Xs  (gpr/model.py:455-483)."""
    with torch.no_grad():
        X, y, Xs = t64(X), t64(y).reshape(-1, 1), t64(Xs)
        Kt = _noisy_gram(kind, p, sigma, X, jitter, data_var)
        Kfs = K(kind, p, X, Xs)                                  # :467
        L = torch.linalg.cholesky(Kt)
        v = torch.linalg.solve_triangular(L, Kfs, upper=False)   # :470
        mu = Kfs.T @ torch.cholesky_solve(y, L)                  # :472
        if full:
            var = K(kind, p, Xs) - v.T @ v                       # :477-478
        else:
            var = (K_diag(kind, p, Xs) - v.T.square().sum(dim=1)).reshape(-1, 1)   # :480-482
        return mu, var


# --------------------------------------------------------------------------------------
# raw-space model: what one reference training iteration costs on the CPU
# --------------------------------------------------------------------------------------

class RawModel:
    """Raw (unconstrained) leaf tensors + reference transforms; ``loss()`` is one
    reference iteration: zero_grad, forward, backward (gpr/model.py:279-292)."""

    def __init__(self, kind, constrained, sigma, X, y, jitter=1e-8, bounds=None):
        self.kind, self.X, self.y, self.jitter = kind, t64(X), t64(y).reshape(-1, 1), jitter
        lo = 1e-8                                               # config.positive_minimum
        self.bounds = {k: (lo, None) for k in constrained}
        if kind == "MOSM":
            self.bounds["delay"] = (None, None)
            self.bounds["phase"] = (None, None)
        if kind == "CONV":
            self.bounds["variance"] = (0.0, None)
        self.bounds["sigma"] = (lo, None)
        if bounds:
            self.bounds.update(bounds)
        self.raw = {}
        allp = dict(constrained); allp["sigma"] = sigma
        for k, v in allp.items():
            v = t64(v).clone()
            lower, upper = self.bounds[k]
            if lower is not None and upper is not None:
                v = sigmoid_inverse(v, lower, t64(upper))
            elif lower is not None:
                v = softplus_inverse(v, lower)
            self.raw[k] = v.requires_grad_(True)

    def constrained(self):
        out = {}
        for k, r in self.raw.items():
            lower, upper = self.bounds[k]
            if lower is not None and upper is not None:
                out[k] = sigmoid_forward(r, lower, t64(upper))
            elif lower is not None:
                out[k] = softplus_forward(r, lower)
            else:
                out[k] = r
        return out

    def loss(self):
        for r in self.raw.values():
            r.grad = None
        c = self.constrained()
        sigma = c.pop("sigma")
        loss = -lml(self.kind, c, sigma, self.X, self.y, self.jitter)
        loss.backward()
        return loss
