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


KSUB = {"CSM": csm_ksub, "SMLMC": smlmc_ksub}
KSUB_DIAG = {"CSM": csm_ksub_diag, "SMLMC": smlmc_ksub_diag}
PARAM_NAMES = {"CSM": ("amplitude", "mean", "variance", "shift"), "SMLMC": ("weight", "magnitude", "mean", "variance")}


def register():
    """Plug both families into oracle.mogp_oracle's block assembly / LML / gradient / prediction restatements."""
    from oracle import mogp_oracle as orc
    orc.register_kind("CSM", PARAM_NAMES["CSM"], csm_ksub, csm_ksub_diag, lambda p: p["amplitude"].shape[1])
    orc.register_kind("SMLMC", PARAM_NAMES["SMLMC"], smlmc_ksub, smlmc_ksub_diag)
    return orc
