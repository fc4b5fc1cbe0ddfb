"""TEST DOUBLE: an Engine look-alike whose arithmetic is the CPU oracle.

It exists only so that the host-side logic of mogptk_b200.gpr / B200Exact (parameter packing,
autograd routing into p.grad, permutation handling, caching, error mapping) can be tested in
the CPU-only container.  The product never uses it: without a GPU the real Engine raises."""
import numpy as np
import torch

from mogptk_b200.engine import PARAM_ORDER, Rows, kernel_dims, param_shapes
from oracle import mogp_oracle as orc
from oracle import next_kernels as _nk

_nk.register()                 # CSM / SM-LMC / uMOSM restatements into the oracle's block assembly


def _ok(kind):
    """Oracle kind name of a host kind ("CSM:2" -> "CSM")."""
    return str(kind).partition(":")[0]


def _unpack(kind, dims, flat):
    C_, Q, D = dims
    out, o = {}, 0
    for name in PARAM_ORDER[kind]:
        shp = param_shapes(kind, C_, Q, D)[name]
        n = int(np.prod(shp))
        out[name] = flat[o:o + n].reshape(shp).clone()
        o += n
    return out


class FakeEngine:
    def __init__(self, max_n=100000):
        self.device = torch.device("cpu")
        self.max_n = max_n
        self._train = None
        self.calls = 0

    def prepare(self, kind, params, X, y, data_var=None):
        C_, Q, D = kernel_dims(kind, params)
        rows = Rows(X, C_, self.device)
        rows.y = rows.sort_vec(y, self.device)
        rows.dv = rows.sort_vec(data_var, self.device) if data_var is not None else None
        rows.dims, rows.kind = (C_, Q, D), kind
        rows.Xs = torch.cat([torch.repeat_interleave(torch.arange(C_, dtype=torch.float64),
                                                     torch.tensor(np.diff(rows.chan_off)))[:, None], rows.x], dim=1)
        return rows

    def lml_grad_prepared(self, rows, packed, sigma, jitter=1e-8, want_grad=True, check=True):
        self.calls += 1
        p = _unpack(rows.kind, rows.dims, packed.detach())
        C_ = rows.dims[0]
        out = torch.zeros(2 + packed.numel() + C_, dtype=torch.float64)
        try:
            if want_grad:
                with torch.enable_grad():      # called from inside autograd.Function.forward (grad mode off)
                    loss, g = orc.loss_and_grad(_ok(rows.kind), p, sigma, rows.Xs, rows.y, jitter, rows.dv)
                out[0] = -loss
                out[2:2 + packed.numel()] = torch.cat([g[n].reshape(-1) for n in PARAM_ORDER[rows.kind]])
                out[2 + packed.numel():] = g["sigma"].reshape(-1)
            else:
                out[0] = orc.lml(_ok(rows.kind), p, sigma, rows.Xs, rows.y, jitter, rows.dv)
        except torch.linalg.LinAlgError:
            out[1] = 1.0
        self._train, self._p, self._sigma, self._jitter = rows, p, sigma.detach().clone(), jitter
        return out

    def alpha(self):
        rows = self._train
        Kn = orc._noisy_gram(_ok(rows.kind), self._p, self._sigma, rows.Xs, self._jitter, rows.dv)
        a = torch.linalg.solve(Kn, rows.y.reshape(-1, 1)).reshape(-1)
        return a if rows.sorted else a[rows.inv_t]

    def predict(self, Xs, full=False):
        rows = self._train
        mu, var = orc.predict_f(_ok(rows.kind), self._p, self._sigma, rows.Xs, rows.y, torch.as_tensor(np.asarray(Xs)),
                                self._jitter, full=full, data_var=rows.dv)
        return mu.reshape(-1), (var if full else var.reshape(-1))

    def K(self, kind, params, X1, X2=None, sigma=None, data_var=None, jitter=0.0):
        p = {k: v.detach() for k, v in params.items()}
        if X2 is None and (sigma is not None or data_var is not None or jitter):
            return orc._noisy_gram(_ok(kind), p, torch.as_tensor(sigma), orc.t64(X1), jitter, data_var)
        return orc.K(_ok(kind), p, X1, X2)

    def K_diag(self, kind, params, X):
        return orc.K_diag(_ok(kind), {k: v.detach() for k, v in params.items()}, X)
