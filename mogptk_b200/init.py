"""Parameter initialisers that run exact GPs, on the engine.

The reference estimates starting values for its multi-output kernels by fitting small single-channel exact GPs:

* ``mogptk.init.BNSE`` (mogptk/init.py:5-126; Tobar 2018): one ``SpectralKernel`` GP per channel and input dimension,
  trained with Adam (lr 2.0), followed by the closed-form posterior of the signal's Fourier transform on a frequency grid;
* ``Data.get_sm_estimation`` (mogptk/data.py:1053-1087): one ``mogptk.SM`` model per channel, trained like any model.

Both construct their GP internally (``gpr.Exact(...)`` / ``SM(self, Q)``), i.e. outside the ``inference=`` seam, so
``mogptk_b200.install()`` swaps in the functions below (same signatures, same return values).  The exact-GP work --
training (device-resident Adam loop), the Gram matrix, its Cholesky factor / inverse and the N^2 n products with the
time-frequency cross-covariances -- goes through the C ABI; only O(N n) element-wise closed forms stay in torch.
Channels are independent small problems: ``bnse_estimation_concurrent`` runs them side by side, each on its own
workspace handle, CUDA stream and host thread.
"""
import math
import threading

import numpy as np
import torch

from . import gpr as _gpr
from .engine import Engine

_tls = threading.local()


def _engine(n_rows, device):
    """One workspace handle per host thread and device (handles are not re-entrant), grown on demand."""
    pool = getattr(_tls, "engines", None)
    if pool is None:
        pool = _tls.engines = {}
    idx = device.index if device.index is not None else torch.cuda.current_device()
    eng = pool.get(idx)
    need = max(256, int(n_rows))
    if eng is None or eng.max_n < need:
        eng = pool[idx] = Engine(device=idx, max_n=need)
    return eng


def _pad128(n):
    return (n + 127) // 128 * 128


def BNSE(x, y, y_err=None, max_freq=None, n=1000, iters=100, jit=True):
    """Bayesian non-parametric spectral estimation: same contract as ``mogptk.init.BNSE`` (mogptk/init.py:5-126).

    Returns (frequencies (n,), PSD mean (n,), PSD variance (n,)) as numpy arrays.  `jit` is accepted and ignored (the
    fused CUDA step has nothing to trace)."""
    dev = _gpr.config.device
    if dev.type != "cuda":
        raise RuntimeError("mogptk_b200.init.BNSE needs a CUDA device: the B200 engine has no CPU fallback")
    x = np.array(x, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    x = x - np.median(x)                                               # init.py:25 (on a copy: the reference shifts in place)
    N = x.shape[0]
    span = float(np.max(x) - np.min(x))
    spacing = span / N
    if max_freq is None:
        max_freq = 0.5 / spacing                                       # init.py:28-29
    max_freq = float(max_freq)

    # ---- the single-channel spectral GP (gpr.SpectralKernel == one SM component of one channel), init.py:39-50
    kernel = _gpr.IndependentMultiOutputKernel([_gpr.SpectralMixtureKernel(Q=1, input_dims=1)], output_dims=1)
    X = np.stack([np.zeros(N), x], axis=1)
    eng = _engine(N, dev)
    sk = kernel[0]
    data_var = None if y_err is None else np.asarray(y_err, dtype=np.float64).reshape(-1) ** 2
    model = _gpr.Exact(kernel, X, y, variance=1.0, data_variance=data_var, engine=eng)
    yt = torch.as_tensor(y)
    sk.magnitude.assign(float(yt.var()))                               # torch's unbiased variance, as the reference uses
    sk.mean.assign(0.01, upper=max_freq)
    sk.variance.assign(0.25 / math.pi ** 2 / spacing ** 2)
    model.likelihood.scale.assign(float(yt.std()) / 10.0)

    # ---- train: Adam, lr = 2.0, `iters` iterations (init.py:55-58: optimizer.step(model.loss))
    if iters > 0:
        from .train import fit_adam
        fit_adam(model, int(iters), lr=2.0, sync_every=64)

    with torch.no_grad():
        mag = sk.magnitude().reshape(()).to(dev)
        mu = sk.mean().reshape(()).to(dev)
        var = sk.variance().reshape(()).to(dev)
        noise = model.likelihood.scale().reshape(()).to(dev)
        alpha = 0.5 / span ** 2                                        # init.py:60
        w = torch.linspace(0.0, max_freq, int(n), device=dev, dtype=torch.float64)          # (n,)
        t = torch.as_tensor(x, device=dev)                                                  # (N,)
        gamma = 2.0 * math.pi ** 2 * var

        # diagonal of the frequency-frequency covariances K(w, w) and K(w, -w) (init.py:63-71; only the diagonal of
        # var_real / var_imag is used at init.py:118-119, so the n x n matrices are never formed here)
        c_ff = 0.5 * math.pi * mag / torch.sqrt(alpha ** 2 + 2.0 * alpha * gamma)
        s2 = 2.0 * math.pi ** 2 / (alpha + 2.0 * gamma)
        k_pp = c_ff * (torch.exp(-s2 * (w - mu) ** 2) + torch.exp(-s2 * (w + mu) ** 2))
        k_pm = c_ff * 2.0 * torch.exp(-0.5 * math.pi ** 2 / alpha * (2.0 * w) ** 2 - s2 * mu ** 2)
        kff_real, kff_imag = 0.5 * (k_pp + k_pm), 0.5 * (k_pp - k_pm)

        # time-frequency cross-covariances, real and imaginary part (init.py:73-92), N x n each
        lq = 1.0 / (math.pi ** 2 * (1.0 / alpha + 1.0 / gamma))
        amp = 0.5 * mag * torch.sqrt(math.pi / (alpha + gamma)) * torch.exp(-math.pi ** 2 * t ** 2 * lq)     # (N,)
        s1 = math.pi ** 2 / (alpha + gamma)
        ea, eb = torch.exp(-s1 * (w - mu) ** 2), torch.exp(-s1 * (w + mu) ** 2)             # (n,)
        tl = (t * lq).reshape(-1, 1)
        pa = -2.0 * math.pi ** 3 * tl * (w / alpha + mu / gamma).reshape(1, -1)             # (N, n)
        pb = -2.0 * math.pi ** 3 * tl * (w / alpha - mu / gamma).reshape(1, -1)
        ktf_real = amp.reshape(-1, 1) * (ea * torch.cos(pa) + eb * torch.cos(pb))
        ktf_imag = amp.reshape(-1, 1) * (ea * torch.sin(pa) + eb * torch.sin(pb))

        # ---- Ktt = K + noise^2 I + jitter * mean(diag) I, its factor and inverse factor on the engine (init.py:95-97)
        kind, p, _ = _gpr.kernel_spec(kernel)
        p = {k: v.detach() for k, v in p.items()}
        Ktt = eng.K(kind, p, X, sigma=noise.reshape(1), jitter=model.jitter)
        Np, npad = _pad128(N), (int(n) + 63) // 64 * 64
        A = torch.eye(Np, dtype=torch.float64, device=dev)
        A[:N, :N] = Ktt
        Linv, _, info = eng.trtri_kinv_(A)                            # A <- L, Linv = L^-1 (lower)
        if info != 0:
            raise _gpr.CholeskyException("linalg.cholesky: the leading minor of order %d is not positive-definite" % info,
                                         Ktt, model)
        Linv = torch.tril(Linv)

        def lsolve(B):                                                 # L^-1 B through the fp64 tensor-pipe GEMM
            Bp = torch.zeros((Np, npad), dtype=torch.float64, device=dev)
            Bp[:N, :B.shape[1]] = B
            out = torch.empty_like(Bp)
            eng.dgemm(0, 0, 1.0, Linv, Bp, 0.0, out)
            return out[:N, :B.shape[1]]

        yv = torch.as_tensor(y, device=dev)
        z = Linv[:N, :N] @ yv                                          # O(N^2) mat-vecs
        a = Linv[:N, :N].T @ z                                         # Ktt^-1 y (init.py:105)
        b, c = lsolve(ktf_real), lsolve(ktf_imag)                      # init.py:106-107
        mu_real, mu_imag = ktf_real.T @ a, ktf_imag.T @ a
        var_real = kff_real - (b * b).sum(dim=0)
        var_imag = kff_imag - (c * c).sum(dim=0)
        # PSD = N(mu_r, var_r)^2 + N(mu_i, var_i)^2: mean and variance of the generalised chi-squared (init.py:116-121)
        psd = mu_real ** 2 + mu_imag ** 2 + var_real + var_imag
        psd_var = 2.0 * var_real ** 2 + 2.0 * var_imag ** 2 + 4.0 * var_real * mu_real ** 2 + 4.0 * var_imag * mu_imag ** 2
        return w.cpu().numpy(), psd.cpu().numpy(), psd_var.cpu().numpy()


def sm_estimation(data, Q=1, method="LS", optimizer="Adam", iters=200, params={}):
    """``Data.get_sm_estimation`` (mogptk/data.py:1053-1087) with the per-channel ``mogptk.SM`` model built through
    ``B200Exact``: returns (amplitudes, means, variances), each (Q, input_dims)."""
    import mogptk

    from .inference import B200Exact
    dims = data.get_input_dims()
    sm = mogptk.SM(data, Q, inference=B200Exact())
    sm.init_parameters(method)
    sm.train(method=optimizer, iters=iters, **params)
    k = sm.gpr.kernel[0]
    return (k.magnitude.numpy().reshape(-1, 1).repeat(dims, axis=1), k.mean.numpy(), k.variance.numpy())


def bnse_estimation_concurrent(dataset, Q=1, n=1000, iters=200):
    """``DataSet.get_bnse_estimation`` (mogptk/dataset.py:605-630) with the channels fitted side by side: every channel
    is an independent small exact GP, so each gets its own host thread, CUDA stream and workspace handle."""
    chans = list(dataset.channels)
    out, errs = [None] * len(chans), []
    dev = _gpr.config.device

    def work(i, ch):
        try:
            stream = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(stream):
                out[i] = ch.get_bnse_estimation(Q, n, iters=iters)
                stream.synchronize()
        except BaseException as e:
            errs.append(e)

    torch.cuda.synchronize()
    threads = [threading.Thread(target=work, args=(i, c)) for i, c in enumerate(chans)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errs:
        raise errs[0]
    return [o[0] for o in out], [o[1] for o in out], [o[2] for o in out]
