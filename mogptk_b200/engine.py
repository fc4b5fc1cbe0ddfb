"""Thin Python host layer over the C ABI: torch CUDA tensors are only the containers
whose ``data_ptr()`` is handed to libmogp_b200.so.  No arithmetic of the hot path
happens here and nothing falls back to the CPU.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi

class _ParamOrder(dict):
    def __missing__(self, kind):                       # "CSM:2" -> the CSM layout
        return self[family(kind)]


PARAM_ORDER = _ParamOrder({
    "MOSM": ("weight", "mean", "variance", "delay", "phase"),
    "SM": ("magnitude", "mean", "variance"),
    "CONV": ("weight", "variance", "base_variance"),
    "CSM": ("amplitude", "mean", "variance", "shift"),
    "SMLMC": ("weight", "magnitude", "mean", "variance"),
    "UMOSM": ("weight", "mean", "variance", "delay", "phase"),
    "MOHSM": ("weight", "mean", "variance", "lengthscale", "center", "delay", "phase"),
})


def family(kind):
    return str(kind).partition(":")[0]


def kind_rq(kind):
    rq = str(kind).partition(":")[2]
    return int(rq) if rq else 1


class NotPositiveDefiniteError(RuntimeError):
    """Cholesky failed: leading minor `info` is not positive definite (LAPACK potrf
    convention).  The host model layer converts this into the reference's
    ``CholeskyException`` (mogptk/gpr/model.py:71-78)."""

    def __init__(self, info):
        super().__init__("linalg.cholesky: the leading minor of order %d is not positive-definite" % info)
        self.info = int(info)


def kernel_dims(kind, params):
    """(C, Q, D) from the constrained parameter shapes."""
    if kind == "MOSM":
        C_, Q, D = params["mean"].shape
    elif kind == "SM":
        C_, Q, D = params["mean"].shape
    elif kind == "CONV":
        Q, C_, D = params["variance"].shape
    elif family(kind) == "CSM":
        Q, C_, _ = params["amplitude"].shape
        D = params["mean"].shape[1]
    elif family(kind) == "SMLMC":
        C_, Q, _ = params["weight"].shape
        D = params["mean"].shape[1]
    elif kind in ("UMOSM", "MOHSM"):
        Q, C_, D = params["mean"].shape
    else:
        raise ValueError("unknown kernel kind %r" % (kind,))
    return int(C_), int(Q), int(D)


def param_shapes(kind, C_, Q, D):
    if kind == "MOSM":
        return {"weight": (C_, Q), "mean": (C_, Q, D), "variance": (C_, Q, D), "delay": (C_, Q, D), "phase": (C_, Q)}
    if kind == "SM":
        return {"magnitude": (C_, Q), "mean": (C_, Q, D), "variance": (C_, Q, D)}
    if kind == "CONV":
        return {"weight": (Q, C_), "variance": (Q, C_, D), "base_variance": (Q, D)}
    Rq = kind_rq(kind)
    if family(kind) == "CSM":
        return {"amplitude": (Q, C_, Rq), "mean": (Q, D), "variance": (Q, D), "shift": (Q, C_, Rq)}
    if family(kind) == "SMLMC":
        return {"weight": (C_, Q, Rq), "magnitude": (Q,), "mean": (Q, D), "variance": (Q, D)}
    if kind == "UMOSM":
        return {"weight": (Q, C_, C_), "mean": (Q, C_, D), "variance": (Q, C_, D), "delay": (Q, C_, D), "phase": (Q, C_)}
    if kind == "MOHSM":
        return {"weight": (Q, C_), "mean": (Q, C_, D), "variance": (Q, C_, D), "lengthscale": (Q, C_), "center": (Q, D),
                "delay": (Q, C_, D), "phase": (Q, C_)}
    raise ValueError("unknown kernel kind %r" % (kind,))


def pack_params(kind, params, device=None):
    """Flatten the constrained parameters into the packed layout of include/mogp_b200.h."""
    parts = []
    for name in PARAM_ORDER[kind]:
        v = params[name]
        if not isinstance(v, torch.Tensor):
            v = torch.as_tensor(np.asarray(v))
        parts.append(v.detach().to(torch.float64).reshape(-1))
    flat = torch.cat(parts)
    return flat.to(device) if device is not None else flat


def unpack_grads(kind, C_, Q, D, flat):
    out, o = {}, 0
    for name in PARAM_ORDER[kind]:
        shp = param_shapes(kind, C_, Q, D)[name]
        n = int(np.prod(shp))
        out[name] = flat[o:o + n].reshape(shp)
        o += n
    return out


class Rows:
    """Channel-sorted view of a kernel-format X (column 0 = channel id)."""

    def __init__(self, X, C_, device):
        X = np.ascontiguousarray(np.asarray(X.detach().cpu() if isinstance(X, torch.Tensor) else X, dtype=np.float64))
        if X.ndim != 2 or X.shape[0] == 0 or X.shape[1] < 2:
            raise ValueError("X must have shape (data_points, 1+input_dims) with channel ids in column 0")
        chan = X[:, 0]
        ids = chan.astype(np.int64)
        if np.any(ids != chan) or np.any(ids < 0) or np.any(ids >= C_):
            raise ValueError("X must have integers for the channel IDs in the first input dimension")
        self.perm = np.argsort(ids, kind="stable")
        self.sorted = bool(np.all(self.perm == np.arange(X.shape[0])))
        counts = np.bincount(ids, minlength=C_)
        self.chan_off = np.zeros(C_ + 1, dtype=np.int32)
        self.chan_off[1:] = np.cumsum(counts)
        self.N = X.shape[0]
        self.D = X.shape[1] - 1
        self.x_host = np.ascontiguousarray(X[self.perm, 1:])
        self.x = torch.from_numpy(self.x_host).to(device)
        self.perm_t = torch.from_numpy(self.perm).to(device)
        self.inv_t = torch.empty_like(self.perm_t)
        self.inv_t[self.perm_t] = torch.arange(self.N, device=device)
        self.off_p = self.chan_off.ctypes.data_as(_cabi.c_ip)

    def sort_vec(self, v, device):
        v = torch.as_tensor(np.asarray(v, dtype=np.float64) if not isinstance(v, torch.Tensor) else v)
        v = v.detach().to(device=device, dtype=torch.float64).reshape(-1)
        return v if self.sorted else v[self.perm_t].contiguous()


class Engine:
    """One workspace handle on one GPU (one process per GPU)."""

    def __init__(self, device=0, max_n=2048):
        if not torch.cuda.is_available():
            raise RuntimeError("mogptk_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _cabi.load()
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        self.max_n = int(max_n)
        h = C.c_void_p()
        rc = self.lib.mogp_create(self.device_index, self.max_n, C.byref(h))
        if rc != 0 or not h:
            raise RuntimeError("mogp_create failed (rc=%d): out of memory or no sm_100a device?" % rc)
        self.h = h
        self._train = None        # Rows of the last lml_grad (for predict)
        self._kind = None
        self._early = None        # numpy view of the handle's early-loss window (False: switched off)

    # ---------------------------------------------------------------- plumbing
    def close(self):
        if getattr(self, "h", None):
            self._early = False                 # (the window lives in memory the handle owns)
            self.lib.mogp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- early loss (see mogp_early_loss in capi.cu)
    def early_loss_buffer(self):
        """The handle's [lml, info, seq] window in mapped pinned host memory (numpy view), or None when switched off
        (MOGP_EARLY_LOSS=0).  Switched on at first use; must be requested BEFORE the evaluation it is to report."""
        if getattr(self, "_early", None) is None:
            import os
            if os.environ.get("MOGP_EARLY_LOSS", "1") == "0":
                self._early = False
            else:
                ptr = C.POINTER(C.c_double)()
                self._check(self.lib.mogp_early_loss(self.h, 1, C.byref(ptr)))
                self._early = np.ctypeslib.as_array(ptr, shape=(3,))
        return self._early if self._early is not False else None

    def wait_early_loss(self, buf, timeout_s=30.0):
        """Spin until the step enqueued last has published its loss; returns (lml, info)."""
        import time
        want = float(self.lib.mogp_early_expected(self.h))
        t0 = None
        while buf[2] != want:
            if t0 is None:
                t0 = time.perf_counter()
            elif time.perf_counter() - t0 > timeout_s:
                torch.cuda.synchronize(self.device)
                if buf[2] != want:
                    raise RuntimeError("early loss: the step did not report (sequence %r, expected %r)" % (float(buf[2]), want))
        return float(buf[0]), float(buf[1])

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.mogp_last_error(self.h)
            raise RuntimeError("libmogp_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def _dev64(self, v):
        if v is None:
            return None
        if not isinstance(v, torch.Tensor):
            v = torch.as_tensor(np.asarray(v, dtype=np.float64))
        return v.detach().to(device=self.device, dtype=torch.float64).contiguous()

    # ---------------------------------------------------------------- kernel matrices
    def K(self, kind, params, X1, X2=None, sigma=None, data_var=None, jitter=0.0):
        """K(X1, X2) (or the Gram matrix with optional noise/jitter diagonal) as an
        (n1, n2) CUDA tensor in the row order of X1 / X2."""
        C_, Q, D = kernel_dims(kind, params)
        r1 = Rows(X1, C_, self.device)
        r2 = Rows(X2, C_, self.device) if X2 is not None else None
        if r1.D != D or (r2 is not None and r2.D != D):
            raise ValueError("input dimensions of X do not match the kernel")
        p = pack_params(kind, params, self.device)
        n2 = r2.N if r2 is not None else r1.N
        out = torch.empty((r1.N, n2), dtype=torch.float64, device=self.device)
        sig = self._dev64(sigma)
        dv = r1.sort_vec(data_var, self.device) if data_var is not None else None
        self._check(self.lib.mogp_kbuild(
            self.h, _cabi.KIND[kind], C_, Q, D, self._p(p), self._p(r1.x), r1.off_p,
            self._p(r2.x) if r2 is not None else C.c_void_p(0), r2.off_p if r2 is not None else None,
            self._p(sig), self._p(dv), float(jitter), self._p(out), n2, self._stream()))
        if not r1.sorted:
            out = out[r1.inv_t]
        if r2 is not None and not r2.sorted:
            out = out[:, r2.inv_t]
        elif r2 is None and not r1.sorted:
            out = out[:, r1.inv_t]
        return out

    def K_diag(self, kind, params, X):
        C_, Q, D = kernel_dims(kind, params)
        r = Rows(X, C_, self.device)
        p = pack_params(kind, params, self.device)
        out = torch.empty(r.N, dtype=torch.float64, device=self.device)
        self._check(self.lib.mogp_kdiag_x(self.h, _cabi.KIND[kind], C_, Q, D, self._p(p), self._p(r.x), r.off_p, self._p(out),
                                          self._stream()))
        return out if r.sorted else out[r.inv_t]

    # ---------------------------------------------------------------- dense building blocks
    def potrf_(self, A):
        """In-place lower Cholesky of a square fp64 CUDA matrix; returns info (int)."""
        assert A.is_cuda and A.dtype == torch.float64 and A.dim() == 2 and A.shape[0] == A.shape[1]
        assert A.stride(1) == 1
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.mogp_potrf(self.h, self._p(A), A.shape[0], A.stride(0), self._p(info), self._stream()))
        return int(info.item())

    def trtri_kinv_(self, A):
        n = A.shape[0]
        Linv = torch.empty_like(A)
        Kinv = torch.empty_like(A)
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._check(self.lib.mogp_trtri_kinv(self.h, self._p(A), self._p(Linv), self._p(Kinv), n, self._p(info),
                                             self._stream()))
        return Linv, Kinv, int(info.item())

    def dgemm(self, transa, transb, alpha, A, B, beta, Cm):
        M, N = Cm.shape
        K = A.shape[0] if transa else A.shape[1]
        self._check(self.lib.mogp_dgemm(self.h, int(transa), int(transb), M, N, K, float(alpha), self._p(A),
                                        A.stride(0), self._p(B), B.stride(0), float(beta), self._p(Cm), Cm.stride(0),
                                        self._stream()))
        return Cm

    def peak_fp64(self):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.mogp_peak_fp64(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---------------------------------------------------------------- the exact-GP step
    def prepare(self, kind, params, X, y, data_var=None):
        """Sort rows by channel and park x / y on the device once (training loops reuse it)."""
        C_, Q, D = kernel_dims(kind, params)
        rows = Rows(X, C_, self.device)
        if rows.D != D:
            raise ValueError("input dimensions of X do not match the kernel")
        if rows.N > self.max_n:
            raise ValueError("N=%d exceeds this engine's max_n=%d" % (rows.N, self.max_n))
        rows.y = rows.sort_vec(y, self.device)
        rows.dv = rows.sort_vec(data_var, self.device) if data_var is not None else None
        rows.dims = (C_, Q, D)
        rows.kind = kind
        return rows

    def lml_grad_prepared(self, rows, packed_params, sigma, jitter=1e-8, want_grad=True, check=True):
        """Device-resident evaluation.  Returns the raw output block (CUDA tensor):
        [lml, info, grad params (P), grad sigma (C)]."""
        C_, Q, D = rows.dims
        P = packed_params.numel()
        out = torch.empty(2 + P + C_, dtype=torch.float64, device=self.device)
        self._check(self.lib.mogp_lml_grad(
            self.h, _cabi.KIND[rows.kind], C_, Q, D, self._p(packed_params), self._p(rows.x), rows.off_p,
            self._p(rows.y), self._p(sigma), self._p(rows.dv), float(jitter), 1 if want_grad else 0, self._p(out),
            self._stream()))
        self._train, self._kind = rows, rows.kind
        if check:
            info = int(out[1].item())
            if info != 0:
                raise NotPositiveDefiniteError(info)
        return out

    def lml_grad(self, kind, params, sigma, X, y, jitter=1e-8, want_grad=True, data_var=None):
        rows = self.prepare(kind, params, X, y, data_var)
        C_, Q, D = rows.dims
        p = pack_params(kind, params, self.device)
        sig = self._dev64(sigma).reshape(-1)
        if sig.numel() == 1 and C_ > 1:
            sig = sig.expand(C_).contiguous()
        out = self.lml_grad_prepared(rows, p, sig, jitter, want_grad)
        res = {"lml": float(out[0].item()), "info": int(out[1].item())}
        if want_grad:
            host = out.cpu()
            res["grad"] = unpack_grads(kind, C_, Q, D, host[2:2 + p.numel()])
            res["grad"]["sigma"] = host[2 + p.numel():]
        return res

    def lml_grad_host(self, kind, dims, packed_params_host, x_host, chan_off, y_host, sigma_host, jitter=1e-8,
                      want_grad=True, out_host=None):
        """End-to-end call with HOST (numpy, ideally pinned) buffers: H2D copies, the step, D2H."""
        C_, Q, D = dims
        P = packed_params_host.size
        if out_host is None:
            out_host = np.empty(2 + P + C_, dtype=np.float64)
        rc = self.lib.mogp_lml_grad_host(
            self.h, _cabi.KIND[kind], C_, Q, D, C.c_void_p(packed_params_host.ctypes.data),
            C.c_void_p(x_host.ctypes.data), chan_off.ctypes.data_as(_cabi.c_ip), C.c_void_p(y_host.ctypes.data),
            C.c_void_p(sigma_host.ctypes.data), C.c_void_p(0), float(jitter), 1 if want_grad else 0,
            C.c_void_p(out_host.ctypes.data))
        self._check(rc)
        return out_host

    def alpha(self):
        """alpha = K~^-1 y of the last evaluation, in the ORIGINAL row order of the training X (d LML / d y = -alpha)."""
        if self._train is None:
            raise RuntimeError("alpha() needs a preceding lml_grad() on this engine")
        rows = self._train
        a = torch.empty(rows.N, dtype=torch.float64, device=self.device)
        self._check(self.lib.mogp_alpha(self.h, self._p(a), self._stream()))
        return a if rows.sorted else a[rows.inv_t]

    def predict(self, Xs, full=False):
        """Posterior mean / variance of f at Xs from the factor of the last lml_grad call."""
        if self._train is None:
            raise RuntimeError("predict() needs a preceding lml_grad() on this engine")
        C_, Q, D = self._train.dims
        if not hasattr(Xs, "shape"):
            Xs = np.asarray(Xs, dtype=np.float64)
        n_s = Xs.shape[0]
        if n_s > self.max_n:                     # the workspace is sized for max_n columns: predict in slices
            if full:
                raise ValueError("full covariance of %d test points exceeds this engine's max_n=%d" % (n_s, self.max_n))
            parts = [self.predict(Xs[i:i + self.max_n]) for i in range(0, n_s, self.max_n)]
            return torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])
        rs = Rows(Xs, C_, self.device)
        if rs.D != D:
            raise ValueError("X must have %d input dimensions" % D)
        M = rs.N
        mu = torch.empty(M, dtype=torch.float64, device=self.device)
        var = torch.empty((M, M) if full else (M,), dtype=torch.float64, device=self.device)
        self._check(self.lib.mogp_predict(self.h, self._p(rs.x), rs.off_p, 1 if full else 0, self._p(mu), self._p(var),
                                          self._stream()))
        if not rs.sorted:
            mu = mu[rs.inv_t]
            var = var[rs.inv_t][:, rs.inv_t] if full else var[rs.inv_t]
        return mu, var
