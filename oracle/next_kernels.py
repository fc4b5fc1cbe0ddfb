"""CPU oracle for the NEXT kernel families on the same hot path.  TEST INFRASTRUCTURE ONLY.

Cross-spectral mixture (CSM) and spectral-mixture LMC (SM-LMC) are not in the product yet; this module restates the
reference's per channel-pair blocks and shows that both fit the one derived form the CUDA kernels evaluate,

    K_ij[a, b] = sum_r alpha_r exp(-1/2 sum_d v_rd u_d^2) cos(2 pi (sum_d m_rd u_d + phi_r)),   u = x_a - x_b + theta_r,

so that adding them is a matter of the per-pair component table (csrc/covmath.cuh) and its chain rule, not of new
kernels.  Pinned against the live reference by oracle/make_golden_next.py (fixtures tests/golden/next_*.npz); checked by
tests/test_next_kernels.py.

Parameter layouts (constrained values, fp64):
  CSM    amplitude (Q,C,Rq)  mean (Q,D)  variance (Q,D)  shift (Q,C,Rq)      MixtureKernel of Q CrossSpectralKernel
  SMLMC  weight (C,Q,Rq)  magnitude (Q,)  mean (Q,D)  variance (Q,D)          LMC of Q SpectralKernel
  UMOSM  weight (Q,C,C) (lower triangle used)  mean/variance/delay (Q,C,D)  phase (Q,C)     mixture of Q uMOSM kernels
  MOHSM  weight (Q,C)  mean/variance/delay (Q,C,D)  lengthscale (Q,C)  center (Q,D)  phase (Q,C)   (non-stationary: one more factor,
         a Gaussian window in the mid-point, and a row-dependent Gram diagonal)
"""
import math

import torch

PI = math.pi


def _tau(x1, x2):
    """Signed difference (n,m,D)  (mogptk/gpr/kernel.py:172-177)."""
    return x1.unsqueeze(1) - x2.unsqueeze(0)


def csm_ksub(i, j, x1, x2, p):
    """Sum over the Q mixture terms (gpr/kernel.py:242-246) of CrossSpectralKernel.Ksub (gpr/multioutput.py:428-449)."""
    tau = _tau(x1, x2)
    out = 0.0
    for q in range(p["amplitude"].shape[0]):
        exp = torch.exp(-0.5 * torch.tensordot(tau ** 2, p["variance"][q], dims=1)).unsqueeze(2)            # :433 / :444
        base = torch.tensordot(tau, p["mean"][q], dims=1).unsqueeze(2)
        if i == j:
            amp = p["amplitude"][q, i].reshape(1, 1, -1)                                                   # :432
            cos = torch.cos(2.0 * PI * base)                                                               # :436
        else:
            shift = p["shift"][q, i] - p["shift"][q, j]                                                    # :439
            amp = torch.sqrt(p["amplitude"][q, i] * p["amplitude"][q, j]).reshape(1, 1, -1)                # :442
            cos = torch.cos(2.0 * PI * (base + shift.reshape(1, 1, -1)))                                   # :447
        out = out + torch.sum(amp * exp * cos, dim=2)
    return out


def smlmc_ksub(i, j, x1, x2, p):
    """LinearModelOfCoregionalizationKernel.Ksub (gpr/multioutput.py:490-495) over SpectralKernel.K
    (gpr/singleoutput.py:550-556; note the einsum SUMS over the input dimensions)."""
    tau = _tau(x1, x2)
    magnitude = torch.sum(p["weight"][i] * p["weight"][j], dim=1)                                          # :493
    out = 0.0
    for q in range(p["magnitude"].shape[0]):
        exp = -2.0 * PI ** 2 * tau ** 2 * p["variance"][q].reshape(1, 1, -1)
        cos = 2.0 * PI * tau * p["mean"][q].reshape(1, 1, -1)
        kq = p["magnitude"][q] * torch.einsum("nmd,nmd->nm", torch.exp(exp), torch.cos(cos))
        out = out + magnitude[q] * kq
    return out


def umosm_ksub(i, j, x1, x2, p):
    """Sum over the Q mixture terms of UncoupledMultiOutputSpectralKernel.Ksub (gpr/multioutput.py:261-286).
    Note the phase sits OUTSIDE the 2 pi factor here (:285), unlike the MOSM kernel (:203)."""
    tau = _tau(x1, x2)
    D = x1.shape[1]
    twopi = (2.0 * PI) ** (D / 2.0)                                                                        # :258
    out = 0.0
    for q in range(p["weight"].shape[0]):
        w = p["weight"][q].tril()
        magnitude = w.mm(w.T)                                                                              # :265
        if i == j:
            variance = p["variance"][q, i]
            alpha = magnitude[i, i] * twopi * variance.prod().sqrt()                                       # :268
            exp = torch.exp(-0.5 * torch.tensordot(tau ** 2, variance, dims=1))
            cos = torch.cos(2.0 * PI * torch.tensordot(tau, p["mean"][q, i], dims=1))
            out = out + alpha * exp * cos
        else:
            iv = 1.0 / (p["variance"][q, i] + p["variance"][q, j])                                         # :273
            dm = p["mean"][q, i] - p["mean"][q, j]
            mag = magnitude[i, j] * torch.exp(-PI ** 2 * dm.dot(iv * dm))                                  # :276
            mean = iv * (p["variance"][q, i] * p["mean"][q, j] + p["variance"][q, j] * p["mean"][q, i])
            variance = 2.0 * p["variance"][q, i] * iv * p["variance"][q, j]
            delay = p["delay"][q, i] - p["delay"][q, j]
            phase = p["phase"][q, i] - p["phase"][q, j]
            alpha = mag * twopi * variance.prod().sqrt()                                                   # :283
            exp = torch.exp(-0.5 * torch.tensordot((tau + delay) ** 2, variance, dims=1))
            cos = torch.cos(2.0 * PI * torch.tensordot(tau + delay, mean, dims=1) + phase)                 # :285
            out = out + alpha * exp * cos
    return out


def umosm_ksub_diag(i, n, p):
    """gpr/multioutput.py:288-293 summed over the mixture."""
    D = p["mean"].shape[2]
    out = 0.0
    for q in range(p["weight"].shape[0]):
        w = p["weight"][q].tril()
        out = out + w.mm(w.T)[i, i] * (2.0 * PI) ** (D / 2.0) * p["variance"][q, i].prod().sqrt()
    return out.repeat(n)


def mohsm_ksub(i, j, x1, x2, p):
    """Sum over the mixture of MultiOutputHarmonizableSpectralKernel.Ksub (gpr/multioutput.py:353-387): the MOSM-like
    stationary factor times a Gaussian window in the mid-point (x + x') / 2 (non-stationary)."""
    tau = _tau(x1, x2)
    avg = 0.5 * (x1.unsqueeze(1) + x2.unsqueeze(0))                                                        # kernel.py average()
    D = x1.shape[1]
    twopi = (2.0 * PI) ** float(D)                                                                         # :350
    ones = torch.ones(D, dtype=torch.float64)
    out = 0.0
    for q in range(p["weight"].shape[0]):
        if i == j:
            variance = p["variance"][q, i]
            ls = p["lengthscale"][q, i] ** 2
            alpha = p["weight"][q, i] ** 2 * twopi * variance.prod().sqrt() * torch.pow(ls.sqrt(), float(D))   # :363
            exp1 = torch.exp(-0.5 * torch.tensordot(tau ** 2, variance, dims=1))
            exp2 = torch.exp(-0.5 * torch.tensordot((avg - p["center"][q]) ** 2, ls * ones, dims=1))
            cos = torch.cos(2.0 * PI * torch.tensordot(tau, p["mean"][q, i], dims=1))
            out = out + alpha * exp1 * cos * exp2
        else:
            li, lj = p["lengthscale"][q, i] ** 2, p["lengthscale"][q, j] ** 2
            iv = 1.0 / (p["variance"][q, i] + p["variance"][q, j])
            il = 1.0 / (li + lj)
            dm = p["mean"][q, i] - p["mean"][q, j]
            mag = p["weight"][q, i] * p["weight"][q, j] * torch.exp(-PI ** 2 * dm.dot(iv * dm))            # :375
            mean = iv * (p["variance"][q, i] * p["mean"][q, j] + p["variance"][q, j] * p["mean"][q, i])
            variance = 2.0 * p["variance"][q, i] * iv * p["variance"][q, j]
            ls = 2.0 * li * il * lj
            delay = p["delay"][q, i] - p["delay"][q, j]
            phase = p["phase"][q, i] - p["phase"][q, j]
            alpha = mag * twopi * variance.prod().sqrt() * torch.pow(ls.sqrt(), float(D))                  # :382
            exp1 = torch.exp(-0.5 * torch.tensordot((tau + delay) ** 2, variance, dims=1))
            exp2 = torch.exp(-0.5 * torch.tensordot((avg - p["center"][q]) ** 2, ls * ones, dims=1))
            cos = torch.cos(2.0 * PI * torch.tensordot(tau + delay, mean, dims=1) + phase)
            out = out + alpha * exp1 * cos * exp2
    return out


def mohsm_ksub_diag(i, x, p):
    """gpr/multioutput.py:389-395 summed over the mixture: the prior variance depends on the input."""
    D = x.shape[1]
    out = 0.0
    for q in range(p["weight"].shape[0]):
        ls = p["lengthscale"][q, i] ** 2
        alpha = p["weight"][q, i] ** 2 * (2.0 * PI) ** float(D) * p["variance"][q, i].prod().sqrt() * torch.pow(ls.sqrt(), float(D))
        out = out + alpha * torch.exp(-0.5 * torch.tensordot((x - p["center"][q]) ** 2, ls * torch.ones(D, dtype=torch.float64), dims=1))
    return out


def derived_components(kind, p, i, j):
    """Per channel-pair component records (alpha, phi, v[D], m[D], theta[D]) of the derived form."""
    comps = []
    if kind == "CSM":
        Q, _, Rq = p["amplitude"].shape
        D = p["mean"].shape[1]
        for q in range(Q):
            for r in range(Rq):
                comps.append((torch.sqrt(p["amplitude"][q, i, r] * p["amplitude"][q, j, r]),
                              p["shift"][q, i, r] - p["shift"][q, j, r],
                              p["variance"][q], p["mean"][q], torch.zeros(D, dtype=torch.float64)))
    elif kind == "SMLMC":
        Q, D = p["mean"].shape
        w = torch.sum(p["weight"][i] * p["weight"][j], dim=1)
        for q in range(Q):
            for d in range(D):                           # one active dimension per component, as for the SM kernel
                e = torch.zeros(D, dtype=torch.float64)
                e[d] = 1.0
                comps.append((w[q] * p["magnitude"][q], torch.zeros((), dtype=torch.float64),
                              4.0 * PI ** 2 * p["variance"][q] * e, p["mean"][q] * e, torch.zeros(D, dtype=torch.float64)))
    elif kind == "UMOSM":
        Q, D = p["weight"].shape[0], p["mean"].shape[2]
        for q in range(Q):
            w = p["weight"][q].tril()
            si, sj, mi, mj = p["variance"][q, i], p["variance"][q, j], p["mean"][q, i], p["mean"][q, j]
            iv = 1.0 / (si + sj)
            dm = mi - mj
            v = 2.0 * si * iv * sj
            alpha = w.mm(w.T)[i, j] * torch.exp(-PI ** 2 * dm.dot(iv * dm)) * (2.0 * PI) ** (D / 2.0) * v.prod().sqrt()
            comps.append((alpha, (p["phase"][q, i] - p["phase"][q, j]) / (2.0 * PI), v, iv * (si * mj + sj * mi),
                          p["delay"][q, i] - p["delay"][q, j]))
    else:
        raise ValueError(kind)
    return comps


def k_from_components(comps, x1, x2):
    out = 0.0
    for alpha, phi, v, m, theta in comps:
        u = _tau(x1, x2) + theta.reshape(1, 1, -1)
        out = out + alpha * torch.exp(-0.5 * torch.tensordot(u ** 2, v, dims=1)) * torch.cos(
            2.0 * PI * (torch.tensordot(u, m, dims=1) + phi))
    return out


def csm_ksub_diag(i, n, p):
    """Sum over the mixture of CrossSpectralKernel.Ksub_diag (gpr/multioutput.py:451-454)."""
    return p["amplitude"][:, i].sum().repeat(n)


def smlmc_ksub_diag(i, n, p):
    """LinearModelOfCoregionalizationKernel.Ksub_diag (gpr/multioutput.py:497-502) over SpectralKernel.K_diag
    (singleoutput.py:558-561: the bare magnitude whatever D is)."""
    magnitude = torch.sum(p["weight"][i] ** 2, dim=1)
    return (magnitude * p["magnitude"]).sum().repeat(n)


KSUB = {"CSM": csm_ksub, "SMLMC": smlmc_ksub, "UMOSM": umosm_ksub, "MOHSM": mohsm_ksub}
KSUB_DIAG = {"CSM": csm_ksub_diag, "SMLMC": smlmc_ksub_diag, "UMOSM": umosm_ksub_diag, "MOHSM": mohsm_ksub_diag}
PARAM_NAMES = {"CSM": ("amplitude", "mean", "variance", "shift"), "SMLMC": ("weight", "magnitude", "mean", "variance"),
               "UMOSM": ("weight", "mean", "variance", "delay", "phase"),
               "MOHSM": ("weight", "mean", "variance", "lengthscale", "center", "delay", "phase")}
# families whose blocks fit the product's derived per-pair component form as it stands (MOHSM needs one more factor: a
# Gaussian window in the mid-point, i.e. a kernel change, not only a table)
DERIVED_FORM = ("CSM", "SMLMC", "UMOSM")


def register():
    """Plug both families into oracle.mogp_oracle's block assembly / LML / gradient / prediction restatements."""
    from oracle import mogp_oracle as orc
    orc.register_kind("CSM", PARAM_NAMES["CSM"], csm_ksub, csm_ksub_diag, lambda p: p["amplitude"].shape[1])
    orc.register_kind("SMLMC", PARAM_NAMES["SMLMC"], smlmc_ksub, smlmc_ksub_diag)
    orc.register_kind("UMOSM", PARAM_NAMES["UMOSM"], umosm_ksub, umosm_ksub_diag, lambda p: p["weight"].shape[1])
    orc.register_kind("MOHSM", PARAM_NAMES["MOHSM"], mohsm_ksub, mohsm_ksub_diag, lambda p: p["weight"].shape[1], diag_needs_x=True)
    return orc
