"""CPU tests for the further kernel families on the hot path (CSM, SM-LMC, uMOSM, MOHSM -- in the product
since round 2; MOHSM is non-stationary: one more factor per component, a Gaussian window in the mid-point): the oracle
restatement reproduces the reference's K (golden fixtures written by oracle/make_golden_next.py from the live reference),
and the per channel-pair component table of the product (csrc/covmath_next.cuh through the host hooks of
libmogp_b200.so) reproduces it too, with an analytic chain rule that matches autograd."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, next_golden_names
from oracle import next_kernels as nk


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    p = {k[2:]: torch.tensor(v, dtype=torch.float64) for k, v in g.items() if k.startswith("p_")}
    X = torch.tensor(g["X"], dtype=torch.float64)
    C = int(g["C"])
    rows = [torch.nonzero(X[:, 0].long() == c, as_tuple=False)[:, 0] for c in range(C)]
    return str(g["kind"]), C, p, X, rows, torch.tensor(g["K"], dtype=torch.float64)


@pytest.mark.parametrize("name", next_golden_names())
def test_restatement_matches_the_reference(name):
    kind, C, p, X, rows, K = _load(name)
    for i in range(C):
        for j in range(C):
            blk = nk.KSUB[kind](i, j, X[rows[i], 1:], X[rows[j], 1:], p)
            assert float((blk - K[rows[i]][:, rows[j]]).abs().max()) <= 1e-13 * float(K.abs().max())


@pytest.mark.parametrize("name", [n for n in next_golden_names() if "mohsm" not in n])
def test_derived_component_form_matches_the_reference(name):
    kind, C, p, X, rows, K = _load(name)
    assert kind in nk.DERIVED_FORM
    for i in range(C):
        for j in range(C):
            comps = nk.derived_components(kind, p, i, j)
            blk = nk.k_from_components(comps, X[rows[i], 1:], X[rows[j], 1:])
            assert float((blk - K[rows[i]][:, rows[j]]).abs().max()) <= 1e-13 * float(K.abs().max())
    # the table is symmetric in the sense the Gram build relies on: block (j, i) is the transpose of block (i, j)
    for i in range(C):
        for j in range(i):
            a = nk.k_from_components(nk.derived_components(kind, p, i, j), X[rows[i], 1:], X[rows[j], 1:])
            b = nk.k_from_components(nk.derived_components(kind, p, j, i), X[rows[j], 1:], X[rows[i], 1:])
            assert float((a - b.T).abs().max()) <= 1e-14 * float(K.abs().max())


def test_fixtures_exist():
    assert len(next_golden_names()) >= 8


# ---------------------------------------------------------------------------------------------------------------
# csrc/covmath_next.cuh through the host hooks of the product library: the component table reproduces the
# reference's K, and the analytic chain rule reproduces autograd through the restatement (a synthetic symmetric
# weight matrix W plays the role of (K^-1 - a a^T) / 2, adj the relative-jitter term on the diagonal pairs).
import ctypes as C

KIND_ID = {"CSM": 3, "SMLMC": 4, "UMOSM": 5, "MOHSM": 6}
ORDER = {"CSM": ("amplitude", "mean", "variance", "shift"), "SMLMC": ("weight", "magnitude", "mean", "variance"),
         "UMOSM": ("weight", "mean", "variance", "delay", "phase"),
         "MOHSM": ("weight", "mean", "variance", "lengthscale", "center", "delay", "phase")}


@pytest.fixture(scope="module")
def explib(lib):
    return lib


def _code(kind, Rq):
    return KIND_ID[kind] | ((Rq << 8) if kind in ("CSM", "SMLMC") else 0)      # MOGP_KIND_WITH_RQ


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _pack(kind, p):
    return np.concatenate([p[k].numpy().reshape(-1) for k in ORDER[kind]]).astype(np.float64)


def _dims(kind, p):
    if kind == "CSM":
        Q, Cn, Rq = p["amplitude"].shape
        return Cn, Q, Rq, p["mean"].shape[1]
    if kind == "UMOSM":
        Q, Cn, _ = p["weight"].shape
        return Cn, Q, 1, p["mean"].shape[2]
    if kind == "MOHSM":
        Q, Cn = p["weight"].shape
        return Cn, Q, 1, p["mean"].shape[2]
    Cn, Q, Rq = p["weight"].shape
    return Cn, Q, Rq, p["mean"].shape[1]


def _block_terms(comp, xa, xb, D, window=False):
    alpha, phi = comp[0], comp[1]
    v, m, th = comp[2:2 + D], comp[2 + D:2 + 2 * D], comp[2 + 2 * D:2 + 3 * D]
    u = xa[:, None, :] - xb[None, :, :] + th[None, None, :]
    E = np.exp(-0.5 * (u ** 2 * v).sum(-1))
    s = None
    if window:                                            # MOHSM: [l, c[D]] after theta; mid-point window
        s = 0.5 * (xa[:, None, :] + xb[None, :, :]) - comp[3 + 3 * D:3 + 4 * D][None, None, :]
        E = E * np.exp(-0.5 * comp[2 + 3 * D] * (s ** 2).sum(-1))
    ang = 2 * np.pi * ((u * m).sum(-1) + phi)
    return alpha, E * np.cos(ang), E * np.sin(ang), u, s


@pytest.mark.parametrize("name", next_golden_names())
def test_component_table_and_chain_rule_of_the_next_families(explib, name):
    kind, Cn, p, X, rows, K = _load(name)
    Cn, Q, Rq, D = _dims(kind, p)
    packed = _pack(kind, p)
    assert explib.mogp_num_params(_code(kind, Rq), Cn, Q, D) == packed.size
    win = kind == "MOHSM"
    st = (3 + 4 * D) if win else (2 + 3 * D)
    R = {"CSM": Q * Rq, "SMLMC": Q * D, "UMOSM": Q, "MOHSM": Q}[kind]
    comps = np.zeros(Cn * Cn * R * st)
    assert explib.mogp_host_pair_comps(_code(kind, Rq), Cn, Q, D, _ptr(packed), _ptr(comps)) == R
    comps = comps.reshape(Cn, Cn, R, st)
    xs = [X[rows[c], 1:].numpy() for c in range(Cn)]
    Kn = K.numpy()
    # 1. the table rebuilds K
    for i in range(Cn):
        for j in range(Cn):
            blk = sum(a * EC for a, EC, _, _, _ in (_block_terms(comps[i, j, r], xs[i], xs[j], D, win) for r in range(R)))
            ref = Kn[np.ix_(rows[i].numpy(), rows[j].numpy())]
            assert np.abs(blk - ref).max() <= 1e-13 * np.abs(Kn).max()
    # 2. chain rule against autograd: loss = sum_ab W_ab K_ab + sum_c adj_c * (diagonal value of pair (c, c))
    rng = np.random.default_rng(7)
    N = X.shape[0]
    order = np.concatenate([rows[c].numpy() for c in range(Cn)])
    off = np.concatenate([[0], np.cumsum([len(rows[c]) for c in range(Cn)])])
    A = rng.standard_normal((N, N))
    W = 0.5 * (A + A.T)                                   # in channel-sorted order
    # (MOHSM: the relative-jitter term of its row-dependent diagonal is folded into gsum by the finalize kernel, adj stays 0)
    adj = np.zeros(Cn) if win else rng.standard_normal(Cn)
    gsum = np.zeros((Cn * (Cn + 1) // 2, R, st))
    for i in range(Cn):
        for j in range(i + 1):
            Wb = W[off[i]:off[i + 1], off[j]:off[j + 1]] * (1.0 if i == j else 2.0)
            for r in range(R):
                _, EC, ES, u, sm = _block_terms(comps[i, j, r], xs[i], xs[j], D, win)
                rec = gsum[i * (i + 1) // 2 + j, r]
                rec[0] = (Wb * EC).sum()
                rec[1] = (Wb * ES).sum()
                for d in range(D):
                    rec[2 + d] = (Wb * EC * u[..., d] ** 2).sum()
                    rec[2 + D + d] = (Wb * ES * u[..., d]).sum()
                    rec[2 + 2 * D + d] = (Wb * EC * u[..., d]).sum()
                    if win:
                        rec[3 + 3 * D + d] = (Wb * EC * sm[..., d]).sum()
                if win:
                    rec[2 + 3 * D] = (Wb * EC * (sm ** 2).sum(-1)).sum()
    grad = np.zeros(packed.size)
    assert explib.mogp_host_chain(_code(kind, Rq), Cn, Q, D, _ptr(packed), _ptr(gsum), _ptr(adj), _ptr(grad)) == packed.size
    pt = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss = 0.0
    xt = [torch.tensor(x) for x in xs]
    Wt = torch.tensor(W)
    for i in range(Cn):
        for j in range(Cn):
            loss = loss + (Wt[off[i]:off[i + 1], off[j]:off[j + 1]] * nk.KSUB[kind](i, j, xt[i], xt[j], pt)).sum()
        zero = torch.zeros(1, D, dtype=torch.float64)
        if not win:
            loss = loss + adj[i] * nk.KSUB[kind](i, i, zero, zero, pt).sum()      # K_rr of channel i (sum of the alphas)
    loss.backward()
    ref = np.concatenate([pt[k].grad.numpy().reshape(-1) for k in ORDER[kind]])
    o = 0
    for k in ORDER[kind]:
        n = pt[k].numel()
        scale = max(np.abs(ref[o:o + n]).max(), 1e-12)
        assert np.abs(grad[o:o + n] - ref[o:o + n]).max() <= 1e-9 * scale, (k, grad[o:o + n], ref[o:o + n])
        o += n


@pytest.mark.parametrize("name", next_golden_names())
def test_oracle_step_of_the_next_families_matches_the_reference(name):
    """LML, gradients w.r.t. the constrained parameters and predictions of the reference's gpr.Exact on these kernels
    (fixtures from the live reference) against the oracle restatement -- the parity target of the GPU path (tests/test_gpu_parity.py::test_further_kernel_families_match_the_reference)."""
    orc = nk.register()
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    kind, C_, p, X, rows, K = _load(name)
    y, sigma, jitter = torch.tensor(z["y"]), torch.tensor(z["sigma"]), float(z["jitter"])
    lml = float(orc.lml(kind, p, sigma, X, y, jitter))
    assert abs(lml - float(z["lml"])) <= 1e-10 * abs(float(z["lml"]))
    _, g = orc.loss_and_grad(kind, p, sigma, X, y, jitter)
    for k in list(nk.PARAM_NAMES[kind]) + ["sigma"]:
        ref = z["gc_" + k]
        assert np.abs(g[k].numpy().reshape(ref.shape) - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-12), k
    mu, var = orc.predict_f(kind, p, sigma, X, y, torch.tensor(z["Xs"]), jitter)
    assert np.abs(mu.numpy().reshape(-1) - z["pred_mu"]).max() <= 1e-9 * max(np.abs(z["pred_mu"]).max(), 1e-12)
    assert np.abs(var.numpy().reshape(-1) - z["pred_var"]).max() <= 1e-9 * max(np.abs(z["pred_var"]).max(), 1e-12)
