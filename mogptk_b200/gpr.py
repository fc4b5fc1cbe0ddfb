"""Host-side mirror of the reference's operator interface for the exact-GP hot path.

Same class names, argument meaning and error behaviour as ``mogptk.gpr`` (GAMES-UChile/mogptk
v0.5.1) for exactly the pieces the path needs -- constrained ``Parameter``s, the MOSM / SM /
CONV kernels as *parameter containers*, the Gaussian likelihood scale, and the ``Exact``
model -- but every O(N^2) / O(N^3) operation goes through libmogp_b200.so (hand-written
sm_100a CUDA behind the C ABI in include/mogp_b200.h).  Nothing here evaluates a kernel
matrix or factorises anything in PyTorch, and there is no CPU fallback: without a GPU the
model raises.

``Exact`` also accepts the reference's *own* kernel objects (duck-typed by class name), which
is how ``mogptk.MOSM(dataset, Q, inference=B200Exact())`` plugs this engine in behind the
reference's builder seam (mogptk/model.py:89-100,231) -- see mogptk_b200/inference.py and
INTEGRATION.md.

Reference lines restated here are cited inline (file:line under /root/reference/mogptk/).
"""
import math
import sys

import numpy as np
import torch

from . import engine as _engine


# --------------------------------------------------------------------------------------
# config  (gpr/config.py:3-10)
# --------------------------------------------------------------------------------------
class Config:
    dtype = torch.float64
    device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    positive_minimum = 1e-8


config = Config()


def use_gpu(n=None):
    """Select CUDA device n (gpr/config.py:51-62)."""
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA is not available")
    if n is None:
        n = torch.cuda.current_device()
    if n < 0 or n >= torch.cuda.device_count():
        raise ValueError("CUDA device %d does not exist" % n)
    torch.cuda.set_device(n)
    config.device = torch.device("cuda", n)


# --------------------------------------------------------------------------------------
# constrained parameters
# --------------------------------------------------------------------------------------
# What the engine needs from a hyper-parameter is small: the unconstrained leaf tensor the optimiser updates, the box
# (lower, upper) that defines its constraint, and which of the three maps of mogp_param_entry (identity / softplus /
# sigmoid, include/mogp_b200.h) turns one into the other.  `Constraint` is that description plus its torch form for
# the autograd path; `Parameter` is a leaf tensor carrying one, with the reference's user-facing behaviour
# (gpr/parameter.py:99-346: p() is the constrained value, assign() stores the unconstrained one, names, pegging,
# priors, pickling) because mogptk.Model and the tests drive it exactly like the reference's class.
_META = ("_name", "lower", "upper", "prior", "train", "pegged_parameter", "pegged_transform", "num_parameters")


def _fit_shape(t, shape, what, scalar_ok=True):
    """Add / drop trailing singleton axes until `t` has `shape` (bounds given as scalars broadcast: scalar_ok)."""
    if t.ndim == 0 and scalar_ok:
        return t
    orig = tuple(t.shape)
    while t.ndim < len(shape) and shape[t.ndim] == 1:
        t = t.unsqueeze(-1)
    while t.ndim > len(shape) and t.shape[-1] == 1:
        t = t.squeeze(-1)
    if tuple(t.shape) != tuple(shape):
        raise ValueError("%s: %s != %s" % (what, orig, tuple(shape)))
    return t


class Constraint:
    """Box constraint of a parameter and the smooth map onto it.

    kind       map (unconstrained x -> constrained y)                         C-ABI entry type
    identity   y = x                                                          0
    softplus   y = b + log(1 + exp(beta x)) / beta, beta = +0.1 (lower bound  1
               b) or -0.1 (upper bound b), linear beyond beta x > 20
    sigmoid    y = lower + (upper - lower) / (1 + exp(-x))                    2
    """

    def __init__(self, lower=None, upper=None):
        self.lower, self.upper = lower, upper
        self.threshold = 20.0
        if lower is not None and upper is not None:
            if torch.any(upper < lower):
                raise ValueError("lower limit %s must be lower than upper limit %s" % (lower, upper))
            self.kind, self.beta, self.offset = "sigmoid", None, lower
        elif lower is not None or upper is not None:
            self.kind = "softplus"
            self.beta, self.offset = (0.1, lower) if lower is not None else (-0.1, upper)
        else:
            self.kind, self.beta, self.offset = "identity", None, None

    def engine_entry(self):
        """(type, beta, lower, upper) of mogp_param_entry."""
        if self.kind == "identity":
            return 0, 0.0, None, None
        if self.kind == "softplus":
            return 1, self.beta, self.offset, None
        return 2, 0.0, self.lower, self.upper

    def forward(self, x):
        if self.kind == "identity":
            return x
        if self.kind == "softplus":
            return self.offset + torch.nn.functional.softplus(x, beta=self.beta, threshold=self.threshold)
        return self.lower + (self.upper - self.lower) * torch.sigmoid(x)

    def inverse(self, y):
        """Unconstrained value stored for a requested constrained value.  For the softplus map this follows the
        reference's expression (gpr/parameter.py:59, `-beta*y - lower` inside the expm1), which is not an exact
        inverse: assign(1.0) reads back 1.0000000995 there and must do so here."""
        if self.kind == "identity":
            return y
        if self.kind == "softplus":
            outside = (self.offset < y) if self.beta < 0.0 else (y < self.offset)
            if torch.any(outside):
                raise ValueError("values must be %s than %s" % ("smaller" if self.beta < 0.0 else "greater", self.offset))
            return (y - self.offset) + torch.log(-torch.expm1(-self.beta * y - self.offset)) / self.beta
        if torch.any(y < self.lower) or torch.any(self.upper < y):
            raise ValueError("values must be between %s and %s" % (self.lower, self.upper))
        frac = (y - self.lower) / (self.upper - self.lower)
        degenerate = torch.isclose(torch.as_tensor(self.lower, dtype=frac.dtype, device=frac.device).expand_as(frac),
                                   torch.as_tensor(self.upper, dtype=frac.dtype, device=frac.device).expand_as(frac))
        frac = torch.where(degenerate, torch.full_like(frac, sys.float_info.epsilon), frac)
        return torch.log(frac) - torch.log1p(-frac)

    def clip(self, y):
        if self.lower is not None:
            y = torch.maximum(y, self.lower.to(y.dtype) * torch.ones_like(y))
        if self.upper is not None:
            y = torch.minimum(y, self.upper.to(y.dtype) * torch.ones_like(y))
        return y


class Parameter(torch.nn.Parameter):
    """Leaf tensor in unconstrained space with its Constraint; ``p()`` / ``p.constrained`` is the constrained value."""

    def __new__(cls, value, name=None, lower=None, upper=None, prior=None, train=True):
        self = super().__new__(cls, Parameter.to_tensor(value))
        self.__dict__.update(_name=name, lower=None, upper=None, prior=prior, train=train, pegged_parameter=None,
                             pegged_transform=None, constraint=Constraint(), num_parameters=int(self.numel()))
        self.assign(self.data, lower=lower, upper=upper)
        return self

    # -- value ---------------------------------------------------------------------------
    @property
    def transform(self):
        """The reference exposes ``transform`` (None when unconstrained); kept for code written against it."""
        return None if self.constraint.kind == "identity" else self.constraint

    @property
    def pegged(self):
        return self.pegged_parameter is not None

    @property
    def constrained(self):
        if self.pegged_parameter is not None:
            src = self.pegged_parameter.constrained
            return src if self.pegged_transform is None else self.pegged_transform(src)
        return self.constraint.forward(self)

    __call__ = lambda self: self.constrained

    def numpy(self):
        return self.constrained.detach().cpu().numpy()

    def __repr__(self):
        vals = self.constrained.tolist()
        return "%s" % (vals,) if self._name is None else "%s=%s" % (self._name, vals)

    def log_prior(self):
        return 0.0 if self.prior is None else self.prior.log_prob(self()).sum()

    @staticmethod
    def to_tensor(value):
        if isinstance(value, Parameter):
            return value.constrained.detach()
        if isinstance(value, torch.Tensor):
            return value.detach().to(config.device, config.dtype)
        return torch.tensor(np.array(value), device=config.device, dtype=config.dtype)

    # -- assignment ----------------------------------------------------------------------
    def _bound(self, b, current, like):
        if b is None:
            return current
        b = b.detach().to(config.device, config.dtype) if torch.is_tensor(b) else \
            torch.tensor(b, device=config.device, dtype=config.dtype)
        return _fit_shape(b, like.shape, "bound and value must match shapes")

    def assign(self, value=None, name=None, lower=None, upper=None, prior=None, train=None):
        """Store the unconstrained image of a constrained `value` and / or change the box, name, prior, train flag.
        Two behaviours of the reference are load-bearing and kept: values outside the box are clipped onto it, and
        changing the bounds WITHOUT a value re-reads the stored unconstrained tensor as if it were the constrained
        value (gpr/parameter.py:232-319; this is what narrows `mean` in mogptk.MOSM's constructor)."""
        if value is None:
            value = self.data
        else:
            value = _fit_shape(Parameter.to_tensor(value), self.shape, "parameter shape must match", scalar_ok=False)
        box = Constraint(self._bound(lower, self.lower, value), self._bound(upper, self.upper, value))
        raw = box.inverse(box.clip(value)) if box.kind != "identity" else value
        raw.requires_grad = True
        self.data = raw
        self.lower, self.upper, self.constraint = box.lower, box.upper, box
        if name is not None:
            head = self._name[:self._name.rfind(".") + 1] if self._name is not None and "." in self._name else ""
            self._name = head + name
        if prior is not None:
            self.prior = prior
        self.train = (True if self.pegged else self.train) if train is None else train
        self.pegged_parameter = self.pegged_transform = None

    def peg(self, other, transform=None):
        if not isinstance(other, Parameter):
            raise ValueError("parameter must be pegged to other parameter object")
        if other.pegged:
            raise ValueError("cannot peg parameter to another pegged parameter")
        self.pegged_parameter, self.pegged_transform, self.train = other, transform, False

    # -- copies / pickles ------------------------------------------------------------------
    def _meta(self):
        return {k: getattr(self, k) for k in _META}

    @staticmethod
    def _restore(data, requires_grad, meta):
        out = torch.nn.Parameter(data, requires_grad)
        out.__class__ = Parameter
        out.__dict__.update(meta)
        out.constraint = Constraint(meta["lower"], meta["upper"])
        return out

    def __deepcopy__(self, memo):
        out = Parameter._restore(self.data.clone(memory_format=torch.preserve_format), self.requires_grad, self._meta())
        memo[id(self)] = out
        return out

    def __reduce_ex__(self, proto):
        return Parameter._restore, (self.data, self.requires_grad, self._meta())


class _Named(torch.nn.Module):
    """Naming / read-only rules shared by kernels, likelihoods and models
    (gpr/kernel.py:37-51, gpr/model.py:138-147)."""

    def __setattr__(self, name, val):
        if name == "train" and not isinstance(self, Exact):
            for p in self.parameters():
                p.train = val
            return
        if hasattr(self, name) and isinstance(getattr(self, name), Parameter):
            raise AttributeError("parameter is read-only, use Parameter.assign()")
        if isinstance(val, Parameter) and val._name is None:
            val._name = "%s.%s" % (self.__class__.__name__, name)
        elif isinstance(val, torch.nn.ModuleList):
            for i, item in enumerate(val):
                for p in item.parameters():
                    p._name = "%s[%d].%s" % (self.__class__.__name__, i, p._name)
        super().__setattr__(name, val)

    def name(self):
        return self.__class__.__name__


# --------------------------------------------------------------------------------------
# kernels as parameter containers; K / K_diag run on the engine
# --------------------------------------------------------------------------------------
_ENGINES = {}


def default_engine(n_rows, device=None):
    """One workspace handle per device, grown on demand."""
    if not torch.cuda.is_available():
        raise RuntimeError("mogptk_b200: no CUDA device -- the B200 engine has no CPU fallback")
    dev = config.device if device is None else device
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    eng = _ENGINES.get(idx)
    if eng is None or eng.max_n < n_rows:
        # a smaller engine may still be referenced by existing models: it is left alone (and released with
        # them); new models get the larger one
        eng = _engine.Engine(device=idx, max_n=max(int(n_rows), 256))
        _ENGINES[idx] = eng
    return eng


def _require_all_dims(kernel, D):
    """The engine evaluates kernels on all D input columns.  In the reference a kernel with ``active_dims`` first
    selects columns of the channel-stripped input (gpr/kernel.py:53-58 under gpr/multioutput.py:26-34,178-181), so
    anything but None / the identity [0..D-1] would be silently ignored here: refuse it."""
    ad = getattr(kernel, "active_dims", None)
    if ad is None:
        return
    vals = ad.detach().cpu().reshape(-1).tolist() if torch.is_tensor(ad) else list(np.asarray(ad).reshape(-1))
    if [int(v) for v in vals] != list(range(D)):
        raise NotImplementedError("active_dims=%s on %s: only all input dimensions are supported by the B200 engine"
                                  % (vals, kernel.__class__.__name__))


def kernel_spec(kernel):
    """Recognise the kernels this engine implements and return (kind, params, (C,Q,D)), where
    params holds the *constrained* tensors (autograd-connected) in the packed order of
    include/mogp_b200.h.  Works for the mirror classes below and for the reference's classes
    of the same name.  Anything else raises NotImplementedError: there is no fallback."""
    cname = kernel.__class__.__name__
    if cname == "MultiOutputSpectralMixtureKernel":
        p = {"weight": kernel.weight(), "mean": kernel.mean(), "variance": kernel.variance(),
             "delay": kernel.delay(), "phase": kernel.phase()}
        C_, Q, D = p["mean"].shape
        _require_all_dims(kernel, int(D))
        return "MOSM", p, (int(C_), int(Q), int(D))
    if cname == "IndependentMultiOutputKernel":
        subs = list(kernel.kernels)
        if all(k.__class__.__name__ == "SpectralMixtureKernel" for k in subs):
            if len({tuple(k.mean.shape) for k in subs}) != 1:
                raise NotImplementedError("SM kernels of all channels must share Q and input_dims")
            p = {"magnitude": torch.stack([k.magnitude() for k in subs]),
                 "mean": torch.stack([k.mean() for k in subs]),
                 "variance": torch.stack([k.variance() for k in subs])}
            C_, Q, D = p["mean"].shape
            for k in [kernel] + subs:
                _require_all_dims(k, int(D))
            return "SM", p, (int(C_), int(Q), int(D))
    if cname in ("MixtureKernel", "AddKernel"):
        subs = list(kernel.kernels)
        if all(k.__class__.__name__ == "CrossSpectralKernel" for k in subs):         # what mogptk.CSM builds
            if len({tuple(k.amplitude.shape) for k in subs}) != 1:
                raise NotImplementedError("CSM terms must share output_dims and Rq")
            p = {"amplitude": torch.stack([k.amplitude() for k in subs]), "mean": torch.stack([k.mean() for k in subs]),
                 "variance": torch.stack([k.variance() for k in subs]), "shift": torch.stack([k.shift() for k in subs])}
            Q, C_, Rq = p["amplitude"].shape
            D = p["mean"].shape[1]
            for k in [kernel] + subs:
                _require_all_dims(k, int(D))
            return "CSM:%d" % Rq, p, (int(C_), int(Q), int(D))
        if all(k.__class__.__name__ == "UncoupledMultiOutputSpectralKernel" for k in subs):
            p = {n: torch.stack([getattr(k, n)() for k in subs]) for n in ("weight", "mean", "variance", "delay", "phase")}
            Q, C_, D = p["mean"].shape
            for k in [kernel] + subs:
                _require_all_dims(k, int(D))
            return "UMOSM", p, (int(C_), int(Q), int(D))
        if all(k.__class__.__name__ == "MultiOutputHarmonizableSpectralKernel" for k in subs):    # what mogptk.MOHSM builds
            names = ("weight", "mean", "variance", "lengthscale", "center", "delay", "phase")
            p = {n: torch.stack([getattr(k, n)() for k in subs]) for n in names}
            Q, C_, D = p["mean"].shape
            for k in [kernel] + subs:
                _require_all_dims(k, int(D))
            return "MOHSM", p, (int(C_), int(Q), int(D))
        if all(k.__class__.__name__ == "GaussianConvolutionProcessKernel" for k in subs):
            p = {"weight": torch.stack([k.weight() for k in subs]),
                 "variance": torch.stack([k.variance() for k in subs]),
                 "base_variance": torch.stack([k.base_variance() for k in subs])}
            Q, C_, D = p["variance"].shape
            for k in [kernel] + subs:
                _require_all_dims(k, int(D))
            return "CONV", p, (int(C_), int(Q), int(D))
    if cname == "LinearModelOfCoregionalizationKernel":                                # what mogptk.SM_LMC builds
        subs = list(kernel.kernels)
        if all(k.__class__.__name__ == "SpectralKernel" for k in subs):
            w = kernel.weight()
            C_, Q, Rq = w.shape
            if Q != len(subs):
                raise NotImplementedError("LMC weight and kernel list disagree on Q")
            p = {"weight": w, "magnitude": torch.stack([k.magnitude().reshape(()) for k in subs]),
                 "mean": torch.stack([k.mean().reshape(-1) for k in subs]),
                 "variance": torch.stack([k.variance().reshape(-1) for k in subs])}
            D = p["mean"].shape[1]
            for k in [kernel] + subs:
                _require_all_dims(k, int(D))
            return "SMLMC:%d" % Rq, p, (int(C_), int(Q), int(D))
    if cname == "GaussianConvolutionProcessKernel":
        p = {"weight": kernel.weight()[None], "variance": kernel.variance()[None],
             "base_variance": kernel.base_variance()[None]}
        Q, C_, D = p["variance"].shape
        _require_all_dims(kernel, int(D))
        return "CONV", p, (int(C_), int(Q), int(D))
    raise NotImplementedError(
        "mogptk_b200 implements the exact-GP path for MOSM, SM (IndependentMultiOutputKernel of SpectralMixtureKernel), "
        "CONV (MixtureKernel of GaussianConvolutionProcessKernel), CSM (MixtureKernel of CrossSpectralKernel), SM-LMC "
        "(LinearModelOfCoregionalizationKernel of SpectralKernel), uMOSM (MixtureKernel of "
        "UncoupledMultiOutputSpectralKernel) and MOHSM (MixtureKernel of MultiOutputHarmonizableSpectralKernel) only; got %s"
        % cname)


def _param_tensors(kind, kernel):
    """The Parameter objects behind kernel_spec's tensors, in the same order (for gradient routing)."""
    fam = _engine.family(kind)
    if fam == "MOSM":
        return [[kernel.weight], [kernel.mean], [kernel.variance], [kernel.delay], [kernel.phase]]
    if fam == "SM":
        subs = list(kernel.kernels)
        return [[k.magnitude for k in subs], [k.mean for k in subs], [k.variance for k in subs]]
    if fam == "SMLMC":
        subs = list(kernel.kernels)
        return [[kernel.weight], [k.magnitude for k in subs], [k.mean for k in subs], [k.variance for k in subs]]
    subs = list(kernel.kernels) if hasattr(kernel, "kernels") else [kernel]
    if fam == "CSM":
        return [[k.amplitude for k in subs], [k.mean for k in subs], [k.variance for k in subs], [k.shift for k in subs]]
    if fam == "UMOSM":
        return [[getattr(k, n) for k in subs] for n in ("weight", "mean", "variance", "delay", "phase")]
    if fam == "MOHSM":
        return [[getattr(k, n) for k in subs] for n in ("weight", "mean", "variance", "lengthscale", "center", "delay", "phase")]
    return [[k.weight for k in subs], [k.variance for k in subs], [k.base_variance for k in subs]]


class Kernel(_Named):
    """Base kernel (gpr/kernel.py:5-191): same call / validation behaviour; K and K_diag are
    computed by the CUDA engine."""

    def __init__(self, input_dims=None, active_dims=None):
        super().__init__()
        self.input_dims = input_dims
        self.active_dims = active_dims
        self.output_dims = None

    def __call__(self, X1, X2=None):
        X1, X2 = self._check_input(X1, X2)
        return self.K(X1, X2)

    def _check_input(self, X1, X2=None):                       # gpr/kernel.py:60-80
        def chk(X):
            if not torch.is_tensor(X):
                X = torch.tensor(np.asarray(X), device=config.device, dtype=config.dtype)
            elif X.device != config.device or X.dtype != config.dtype:
                X = X.to(config.device, config.dtype)
            if X.ndim != 2:
                raise ValueError("X should have two dimensions (data_points,input_dims)")
            return X
        X1 = chk(X1)
        if X1.shape[0] == 0 or X1.shape[1] == 0:
            raise ValueError("X must not be empty")
        if X2 is not None:
            X2 = chk(X2)
            if X2.shape[0] == 0:
                raise ValueError("X must not be empty")
            if X1.shape[1] != X2.shape[1]:
                raise ValueError("input dimensions for X1 and X2 must match")
        return X1, X2

    def iterkernels(self):
        yield self

    def K(self, X1, X2=None):
        kind, p, _ = kernel_spec(self)
        n = X1.shape[0] if X2 is None else max(X1.shape[0], X2.shape[0])
        with torch.no_grad():
            return default_engine(n).K(kind, p, X1, X2)

    def K_diag(self, X1):
        kind, p, _ = kernel_spec(self)
        with torch.no_grad():
            return default_engine(X1.shape[0]).K_diag(kind, p, X1)


class MultiOutputKernel(Kernel):
    """Channel ids live in column 0 of X (gpr/kernel.py:381-404)."""

    def __init__(self, output_dims, input_dims=None, active_dims=None):
        super().__init__(input_dims, active_dims)
        self.output_dims = output_dims

    def _check_input(self, X1, X2=None):
        X1, X2 = super()._check_input(X1, X2)
        for X in (X1, X2):
            if X is None:
                continue
            if not torch.all(X[:, 0] == X[:, 0].long()) or not torch.all(X[:, 0] < self.output_dims):
                raise ValueError("X must have integers for the channel IDs in the first input dimension")
        return X1, X2


class MultiOutputSpectralMixtureKernel(MultiOutputKernel):
    """MOSM parameters (gpr/multioutput.py:156-176): weight (C,Q), mean/variance/delay (C,Q,D), phase (C,Q)."""

    def __init__(self, Q, output_dims, input_dims=1, active_dims=None):
        super().__init__(output_dims, input_dims, active_dims)
        self.input_dims = input_dims
        self.weight = Parameter(torch.ones(output_dims, Q), lower=config.positive_minimum)
        self.mean = Parameter(torch.zeros(output_dims, Q, input_dims), lower=config.positive_minimum)
        self.variance = Parameter(torch.ones(output_dims, Q, input_dims), lower=config.positive_minimum)
        self.delay = Parameter(torch.zeros(output_dims, Q, input_dims))
        self.phase = Parameter(torch.zeros(output_dims, Q))
        if output_dims == 1:
            self.delay.train = False
            self.phase.train = False


class SpectralMixtureKernel(Kernel):
    """SM parameters (gpr/singleoutput.py:583-592): magnitude (Q,), mean/variance (Q,D)."""

    def __init__(self, Q=1, input_dims=1, active_dims=None):
        super().__init__(input_dims, active_dims)
        self.magnitude = Parameter(torch.ones(Q), lower=config.positive_minimum)
        self.mean = Parameter(torch.zeros(Q, input_dims), lower=config.positive_minimum)
        self.variance = Parameter(torch.ones(Q, input_dims), lower=config.positive_minimum)

    def K(self, X1, X2=None):
        raise NotImplementedError("use IndependentMultiOutputKernel([SpectralMixtureKernel...]) (what mogptk.SM builds)")


def _kernel_list(kernels, length=None):
    """gpr/kernel.py:82-110 (argument normalisation and checks)."""
    if isinstance(kernels, tuple):
        kernels = kernels[0] if len(kernels) == 1 and isinstance(kernels[0], list) else list(kernels)
    elif not isinstance(kernels, list):
        kernels = [kernels]
    if len(kernels) == 0:
        raise ValueError("must pass at least one kernel")
    if length is not None and len(kernels) != length:
        if len(kernels) != 1:
            raise ValueError("must pass %d kernels" % length)
        import copy
        kernels = kernels + [copy.deepcopy(kernels[0]) for _ in range(length - 1)]
    for k in kernels:
        if not isinstance(k, Kernel):
            raise ValueError("must pass kernels")
    if any(k.input_dims != kernels[0].input_dims for k in kernels[1:]):
        raise ValueError("kernels must have same input dimensions")
    return kernels


class IndependentMultiOutputKernel(MultiOutputKernel):
    """One sub-kernel per channel, zero cross-covariance (gpr/multioutput.py:5-39)."""

    def __init__(self, *kernels, output_dims=None):
        if output_dims is None:
            output_dims = len(kernels[0]) if len(kernels) == 1 and isinstance(kernels[0], list) else len(kernels)
        super().__init__(output_dims)
        self.kernels = torch.nn.ModuleList(_kernel_list(kernels, output_dims))
        self.input_dims = self.kernels[0].input_dims

    def __getitem__(self, key):
        return self.kernels[key]

    def name(self):
        return "%s[%s]" % (self.__class__.__name__, ",".join(k.name() for k in self.kernels))

    def iterkernels(self):
        yield self
        for k in self.kernels:
            yield k


class GaussianConvolutionProcessKernel(MultiOutputKernel):
    """CONV parameters (gpr/multioutput.py:520-529): weight (C,), variance (C,D) >= 0, base_variance (D,)."""

    def __init__(self, output_dims, input_dims=1, active_dims=None):
        super().__init__(output_dims, input_dims, active_dims)
        self.weight = Parameter(torch.ones(output_dims), lower=config.positive_minimum)
        self.variance = Parameter(torch.ones(output_dims, input_dims), lower=0.0)
        self.base_variance = Parameter(torch.ones(input_dims), lower=config.positive_minimum)


class CrossSpectralKernel(MultiOutputKernel):
    """CSM term (gpr/multioutput.py:397-426): amplitude / shift (C,Rq), mean / variance (D,)."""

    def __init__(self, output_dims, input_dims=1, Rq=1, active_dims=None):
        super().__init__(output_dims, input_dims, active_dims)
        self.amplitude = Parameter(torch.ones(output_dims, Rq), lower=config.positive_minimum)
        self.mean = Parameter(torch.zeros(input_dims), lower=config.positive_minimum)
        self.variance = Parameter(torch.ones(input_dims), lower=config.positive_minimum)
        self.shift = Parameter(torch.zeros(output_dims, Rq))


class UncoupledMultiOutputSpectralKernel(MultiOutputKernel):
    """uMOSM term (gpr/multioutput.py:212-259): weight (C,C) lower triangle, mean / variance / delay (C,D), phase (C,)."""

    def __init__(self, output_dims, input_dims=1, active_dims=None):
        super().__init__(output_dims, input_dims, active_dims)
        self.weight = Parameter(torch.ones(output_dims, output_dims).tril())
        self.weight.num_parameters = (output_dims * output_dims + output_dims) // 2
        self.mean = Parameter(torch.zeros(output_dims, input_dims), lower=config.positive_minimum)
        self.variance = Parameter(torch.ones(output_dims, input_dims), lower=config.positive_minimum)
        self.delay = Parameter(torch.zeros(output_dims, input_dims))
        self.phase = Parameter(torch.zeros(output_dims))
        if output_dims == 1:
            self.delay.train = False
            self.phase.train = False


class MultiOutputHarmonizableSpectralKernel(MultiOutputKernel):
    """MOHSM term (gpr/multioutput.py:295-351): weight / lengthscale / phase (C,), mean / variance / delay (C,D), center (D,).
    Non-stationary: a Gaussian window in the mid-point of the two inputs multiplies the MOSM-like stationary factor."""

    def __init__(self, output_dims, input_dims=1, active_dims=None):
        super().__init__(output_dims, input_dims, active_dims)
        self.weight = Parameter(torch.ones(output_dims), lower=config.positive_minimum)
        self.mean = Parameter(torch.zeros(output_dims, input_dims), lower=config.positive_minimum)
        self.variance = Parameter(torch.ones(output_dims, input_dims), lower=config.positive_minimum)
        self.lengthscale = Parameter(torch.ones(output_dims), lower=config.positive_minimum)
        self.center = Parameter(torch.zeros(input_dims))
        self.delay = Parameter(torch.zeros(output_dims, input_dims))
        self.phase = Parameter(torch.zeros(output_dims))
        if output_dims == 1:
            self.delay.train = False
            self.phase.train = False


class SpectralKernel(Kernel):
    """Single spectral term (gpr/singleoutput.py:520-548): magnitude (), mean / variance (D,)."""

    def __init__(self, input_dims=1, active_dims=None):
        super().__init__(input_dims, active_dims)
        self.magnitude = Parameter(1.0, lower=config.positive_minimum)
        self.mean = Parameter(torch.zeros(input_dims), lower=config.positive_minimum)
        self.variance = Parameter(torch.ones(input_dims), lower=config.positive_minimum)

    def K(self, X1, X2=None):
        raise NotImplementedError("use LinearModelOfCoregionalizationKernel([SpectralKernel...]) (what mogptk.SM_LMC builds)")


class LinearModelOfCoregionalizationKernel(MultiOutputKernel):
    """LMC over Q single-output kernels (gpr/multioutput.py:456-488): weight (C,Q,Rq)."""

    def __init__(self, *kernels, output_dims, input_dims=1, Q=None, Rq=1):
        super().__init__(output_dims, input_dims)
        if Q is None:
            Q = len(kernels[0]) if len(kernels) == 1 and isinstance(kernels[0], list) else len(kernels)
        self.kernels = torch.nn.ModuleList(_kernel_list(kernels, Q))
        self.weight = Parameter(torch.ones(output_dims, Q, Rq), lower=config.positive_minimum)

    def __getitem__(self, key):
        return self.kernels[key]

    def name(self):
        return "%s[%s]" % (self.__class__.__name__, ",".join(k.name() for k in self.kernels))


class AddKernel(Kernel):
    """Sum of kernels (gpr/kernel.py:232-246); sums of CONV / CSM / uMOSM terms reach the engine."""

    def __init__(self, *kernels):
        super().__init__()
        ks = _kernel_list(kernels)
        self.kernels = torch.nn.ModuleList(ks)
        self.input_dims = ks[0].input_dims
        outs = [k.output_dims for k in ks if k.output_dims is not None]
        self.output_dims = outs[0] if outs else None

    def __getitem__(self, key):
        return self.kernels[key]

    def name(self):
        return "[%s]" % ",".join(k.name() for k in self.kernels)

    def iterkernels(self):
        yield self
        for k in self.kernels:
            yield k

    def _check_input(self, X1, X2=None):
        return self.kernels[0]._check_input(X1, X2)


class MixtureKernel(AddKernel):
    """Q copies of one kernel, summed (gpr/kernel.py:264-276)."""

    def __init__(self, kernel, Q):
        if not isinstance(kernel, Kernel):
            raise ValueError("must pass kernel")
        super().__init__(*_kernel_list(kernel, Q))


# --------------------------------------------------------------------------------------
# likelihood  (gpr/likelihood.py:312-378)
# --------------------------------------------------------------------------------------
class GaussianLikelihood(_Named):
    def __init__(self, scale=1.0):
        super().__init__()
        self.output_dims = None
        self.scale = Parameter(scale, lower=config.positive_minimum)
        if self.scale.ndim == 1:
            self.output_dims = self.scale.shape[0]

    def validate_y(self, X, y):
        pass

    def conditional_sample(self, X, f):
        scale = self.scale()
        if self.output_dims is not None:
            scale = scale[X[:, 0].long()].reshape(-1, *([1] * (f.ndim - 1)))
        return torch.distributions.normal.Normal(f, scale=scale).sample()

    def predict(self, X, mu, var, ci=None, sigma=None, n=10000):
        """Confidence band (gpr/likelihood.py:351-378).  For a per-channel scale the reference
        replaces the predictive variance by scale^2 (:355-356): mirrored."""
        if ci is None and sigma is None:
            return mu
        if self.output_dims is not None:
            scale = self.scale()[X[:, 0].long()].reshape(-1, 1)
            if sigma is None:
                c = torch.tensor(ci, device=config.device, dtype=config.dtype)
                lower = mu + math.sqrt(2.0) * scale * torch.erfinv(2.0 * c[0] - 1.0)
                upper = mu + math.sqrt(2.0) * scale * torch.erfinv(2.0 * c[1] - 1.0)
            else:
                lower, upper = mu - sigma * scale, mu + sigma * scale
            return mu, lower, upper
        var = var + self.scale() ** 2
        if sigma is None:
            c = torch.tensor(ci, device=config.device, dtype=config.dtype)
            lower = mu + torch.sqrt(2.0 * var) * torch.erfinv(2.0 * c[0] - 1.0)
            upper = mu + torch.sqrt(2.0 * var) * torch.erfinv(2.0 * c[1] - 1.0)
        else:
            lower, upper = mu - sigma * var.sqrt(), mu + sigma * var.sqrt()
        return mu, lower, upper


# --------------------------------------------------------------------------------------
# the model  (gpr/model.py:71-483)
# --------------------------------------------------------------------------------------
class CholeskyException(Exception):
    """gpr/model.py:71-78.  (When the reference is importable, ``Exact`` raises the reference's
    own class instead so that ``except mogptk.CholeskyException`` keeps working.)"""

    def __init__(self, message, K, model):
        self.message, self.K, self.model = message, K, model

    def __str__(self):
        return self.message


class _ExactLML(torch.autograd.Function):
    """log p(y) with the analytic gradient from the CUDA engine: forward = one fused
    mogp_lml_grad call (K build, Cholesky, inverse, LML, gradient), backward = a scale.

    ``targets`` (y - mean(X), original row order) is an input only when a mean function is present: the reference
    back-propagates through y - mean(X) into the mean's parameters (gpr/model.py:445-452), and
    d LML / d targets = -K~^-1 targets = -alpha, which the engine already holds (mogp_alpha)."""

    @staticmethod
    def forward(ctx, packed, sigma, targets, model):
        out = model._evaluate(packed.detach(), sigma.detach(), want_grad=True,
                              targets=None if targets is None else targets.detach())
        P = packed.numel()
        alpha = None
        if targets is not None and targets.requires_grad:
            alpha = model._eng().alpha().to(targets.device).reshape(targets.shape)
        ctx.has_targets = alpha is not None
        saved = [out[2:2 + P].to(packed.device), out[2 + P:2 + P + sigma.numel()].to(sigma.device)]
        if alpha is not None:
            saved.append(alpha)
        ctx.save_for_backward(*saved)
        return out[0].clone().to(packed.device)

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors           # d(-LML)/d constrained [, alpha]
        gy = -g * saved[2] if ctx.has_targets else None
        return -g * saved[0], -g * saved[1], gy, None


class Exact(_Named):
    """Exact GP regression with a Gaussian likelihood (gpr/model.py:403-483) on the B200 engine.

    Same constructor arguments, attributes (kernel, likelihood, mean, X, y, jitter, input_dims)
    and methods as the reference model, so ``mogptk.Model`` can drive it through
    ``loss()/parameters()/predict_y()/K()/sample_y()``."""

    def __init__(self, kernel, X, y, variance=1.0, data_variance=None, jitter=1e-8, mean=None, engine=None,
                 likelihood_cls=None, cholesky_exception=None):
        super().__init__()
        self._kind, _, self._dims = kernel_spec(kernel)       # raises NotImplementedError for other kernels
        X, y = self._check_input(X, y)
        if mean is not None:
            mu = mean(X).reshape(-1, 1)
            if mu.shape != y.shape:
                raise ValueError("mean and y data must match shapes: %s != %s" % (mu.shape, y.shape))
        variance = Parameter.to_tensor(variance)
        channels = kernel.output_dims if kernel.output_dims is not None else 1
        if 1 < variance.ndim or variance.ndim == 1 and variance.shape[0] != channels:
            raise ValueError("variance must be float or have shape (channels,)")
        if data_variance is not None:
            data_variance = Parameter.to_tensor(data_variance)
            if data_variance.ndim != 1 or data_variance.shape[0] != X.shape[0]:
                raise ValueError("data variance must have shape (data_points,)")
        if config.dtype != torch.float64:
            raise NotImplementedError("the B200 engine computes in float64 only")
        self.data_variance = data_variance
        self.kernel = kernel
        self.X, self.y, self.mean = X, y, mean
        self.likelihood = (likelihood_cls or GaussianLikelihood)(torch.sqrt(variance))
        self.jitter = max(jitter, 1e-15)                       # gpr/model.py:107-110
        self.input_dims = X.shape[1]
        self._compiled_forward = None
        self._engine = engine
        self._rows = None
        self._factor_key = None
        self._chol_exc = cholesky_exception or CholeskyException
        self.log_marginal_likelihood_constant = 0.5 * X.shape[0] * math.log(2.0 * math.pi)

    # ---- reference surface ----------------------------------------------------------
    def name(self):
        return "Exact"

    def _get_name(self):
        return "Exact"

    def compile(self):
        """The reference traces forward() with torch.jit (gpr/model.py:127-129); the fused CUDA
        step has nothing to trace, so this is a no-op kept for API compatibility."""
        self._compiled_forward = None

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None          # device handles are not pickled (recreated lazily)
        state["_rows"] = None
        state["_factor_key"] = None
        state.pop("_fast_cache", None)
        return state

    def _check_input(self, X, y=None):                          # gpr/model.py:149-181
        X = torch.as_tensor(np.asarray(X) if not torch.is_tensor(X) else X).to(config.device, config.dtype)
        if X.ndim == 0:
            X = X.reshape(1, 1)
        elif X.ndim == 1:
            X = X.reshape(-1, 1)
        elif X.ndim != 2:
            raise ValueError("X must have dimensions (data_points,input_dims) with input_dims optional")
        if X.shape[0] == 0 or X.shape[1] == 0:
            raise ValueError("X must not be empty")
        if y is None:
            if X.shape[1] != self.input_dims:
                raise ValueError("X must have %s input dimensions" % self.input_dims)
            return X
        y = torch.as_tensor(np.asarray(y) if not torch.is_tensor(y) else y).to(config.device, config.dtype)
        if y.ndim == 0:
            y = y.reshape(1, 1)
        elif y.ndim == 1:
            y = y.reshape(-1, 1)
        elif y.ndim != 2 or y.shape[1] != 1:
            raise ValueError("y must have one dimension (data_points,)")
        if X.shape[0] != y.shape[0]:
            raise ValueError("number of data points for X and y must match")
        return X, y

    def print_parameters(self, file=None):
        rows = [(p._name, p.numpy().tolist()) for p in self.parameters()]
        width = max([len(r[0]) for r in rows] + [4])
        print("%-*s  %s" % (width, "Name", "Value"), file=file)
        for n, v in rows:
            print("%-*s  %s" % (width, n, v), file=file)

    def log_prior(self):
        return sum(p.log_prior() for p in self.parameters())

    def forward(self, x=None):
        return -self.log_marginal_likelihood() - self.log_prior()

    def loss(self):
        """zero_grad, forward, backward -- one training iteration's evaluation (gpr/model.py:279-292).

        When nothing needs torch's autograd (no mean function, priors or pegged parameters) the whole
        iteration stays on the device: raw leaves -> constrained values (mogp_params_forward), the fused
        exact-GP step (mogp_lml_grad, a replayed CUDA graph for small problems), chain rule into the
        ``p.grad`` buffers (mogp_params_backward) -- three C calls and one synchronisation (the reference
        also synchronises once per iteration on float(loss), mogptk/model.py:384)."""
        fast = self._fast_table()
        if fast is None:
            self.zero_grad(set_to_none=True)
            loss = self.forward()
            loss.backward()
            return loss
        return self._fast_loss(*fast)

    # ---- device-resident iteration --------------------------------
    def _fast_table(self):
        """Entry table for mogp_params_forward/backward, or None if the general autograd path is needed."""
        import ctypes as C
        from ._cabi import ParamEntry
        if self.mean is not None or not torch.cuda.is_available():
            return None
        eng = self._eng()
        if not hasattr(eng, "lib"):
            return None                                        # test double
        kind = self._kind
        groups = _param_tensors(kind, self.kernel) + [[self.likelihood.scale]]
        plist = [q for grp in groups for q in grp]
        C_ = self._dims[0]
        cache = self.__dict__.setdefault("_fast_cache", {})
        for q in plist:
            if getattr(q, "prior", None) is not None or getattr(q, "pegged", False):
                return None
            if not (q.is_cuda and q.dtype == torch.float64 and q.is_contiguous() and q.device == eng.device):
                return None
        if self.likelihood.scale.numel() != C_:
            return None
        sig = tuple((id(q), q.data_ptr(), id(q.transform)) for q in plist)
        if cache.get("sig") == sig:
            for q, gb in zip(plist, cache["grads"]):
                if q.grad is not gb:
                    q.grad = gb
            return cache["entries"], len(plist), cache["P"], plist
        entries = (ParamEntry * len(plist))()
        keep, grads, off = [], [], 0

        def dev(v, like):
            t = v if torch.is_tensor(v) else torch.tensor(float(v))
            t = t.detach().to(device=eng.device, dtype=torch.float64)
            if t.numel() != 1:
                t = t.expand_as(like)
            return t.reshape(-1).contiguous()

        for i, q in enumerate(plist):
            t = q.transform
            e = entries[i]
            e.raw, e.n, e.off = q.data_ptr(), q.numel(), off
            gb = torch.zeros_like(q.data)
            q.grad = gb
            grads.append(gb)
            e.grad = gb.data_ptr()
            e.lower = e.upper = None
            e.lower_n = e.upper_n = 0
            e.type, e.beta = 0, 0.0
            if t is not None:
                if hasattr(t, "engine_entry"):                   # mogptk_b200.gpr.Constraint
                    typ, beta, lo_b, up_b = t.engine_entry()
                else:                                            # the reference's transform objects (gpr/parameter.py:30-96)
                    name = t.__class__.__name__
                    if name == "Softplus" and getattr(t, "threshold", 20.0) == 20.0:
                        typ, beta, lo_b, up_b = 1, float(t.beta), t.lower, None
                    elif name == "Sigmoid":
                        typ, beta, lo_b, up_b = 2, 0.0, t.lower, t.upper
                    else:
                        return None
                e.type, e.beta = typ, float(beta)
                if lo_b is not None:
                    lo = dev(lo_b, q.data)
                    keep.append(lo)
                    e.lower, e.lower_n = lo.data_ptr(), lo.numel()
                if up_b is not None:
                    up = dev(up_b, q.data)
                    keep.append(up)
                    e.upper, e.upper_n = up.data_ptr(), up.numel()
            off += q.numel()
        P = off - C_
        n = 2 + P + C_
        bufs = torch.zeros(3 * n + 8, dtype=torch.float64, device=eng.device)
        cache.update(sig=sig, entries=entries, keep=keep, grads=grads, P=P, bufs=bufs)
        return entries, len(plist), P, plist

    def _fast_loss(self, entries, n_entries, P, plist):
        import ctypes as C
        from . import _cabi
        eng = self._eng()
        lib = eng.lib
        C_, Q, D = self._dims
        cache = self._fast_cache
        bufs = cache["bufs"]
        n = 2 + P + C_
        packed, dcons, out, lossb = bufs[:n], bufs[n:2 * n], bufs[2 * n:3 * n], bufs[3 * n:3 * n + 1]
        if self._rows is None or self._rows.owner is not eng:
            kind, p, _ = kernel_spec(self.kernel)
            rows = eng.prepare(kind, {k: v.detach() for k, v in p.items()}, self.X, self._targets().detach(),
                               self.data_variance)
            rows.owner = eng
            self._rows = rows
        rows = self._rows
        st = eng._stream()
        base = bufs.data_ptr()
        early = eng.early_loss_buffer()          # (before the call: the step must know that it has to report early)
        eng._check(lib.mogp_loss_grad(eng.h, _cabi.KIND[self._kind], C_, Q, D, C.addressof(entries), n_entries, eng._p(rows.x),
                                      rows.off_p, eng._p(rows.y), eng._p(rows.dv), float(self.jitter), C.c_void_p(base),
                                      C.c_void_p(base + 24 * n), st))
        eng._train, eng._kind = rows, self._kind
        if early is not None:
            # the step writes [lml, info, sequence number] into mapped pinned memory right after the solves: return as soon as the
            # loss is known.  K^-1, the gradient reduction and the chain rule into p.grad are still running then; whatever
            # comes next on this stream (the optimiser's kernels) is ordered behind them, as for any asynchronous CUDA op.
            lml, info = eng.wait_early_loss(early)
        else:
            host = cache.get("host")
            if host is None:
                host = cache["host"] = torch.zeros(2, dtype=torch.float64).pin_memory()
            host.copy_(out[:2], non_blocking=True)                 # [lml, info] into pinned memory:
            torch.cuda.current_stream(eng.device).synchronize()    # the iteration's one synchronisation
            lml, info = float(host[0]), float(host[1])
        if info != 0:
            # the reference raises in forward(), before backward() has produced anything (gpr/model.py:246-255,291):
            # do not leave the non-finite chain-rule output in p.grad for callers that catch and continue
            for q in plist:
                q.grad = None
            self._raise_cholesky(int(info), packed[:P], packed[P:P + C_])
        self._factor_key = ("v", tuple((q.data_ptr(), q._version) for q in plist))
        # a host scalar: the training loop reads float(loss) next (mogptk/model.py:384); nothing downstream differentiates it
        return torch.tensor(-lml, dtype=torch.float64)

    def fit_adam(self, iters, **kwargs):
        """`iters` Adam iterations on the device, one host synchronisation per chunk (mogptk_b200.train.fit_adam);
        returns (losses, times)."""
        from .train import fit_adam
        return fit_adam(self, iters, **kwargs)

    def _raise_cholesky(self, info, packed, sigma):
        eng = self._eng()
        kind, p, _ = kernel_spec(self.kernel)
        msg = "linalg.cholesky: the leading minor of order %d is not positive-definite" % info
        with torch.no_grad():
            K = eng.K(kind, {k: v.detach() for k, v in p.items()}, self.X, sigma=sigma, data_var=self.data_variance,
                      jitter=self.jitter)
        print("ERROR:", msg, file=sys.__stdout__)
        if K.isnan().any():
            print("ERROR: kernel matrix has NaNs!", file=sys.__stdout__)
        if K.isinf().any():
            print("ERROR: kernel matrix has infinities!", file=sys.__stdout__)
        raise self._chol_exc(msg, K, self)

    # ---- engine plumbing -------------------------------------------------------------
    def _eng(self):
        if self._engine is None:
            self._engine = default_engine(self.X.shape[0], self.X.device)
        return self._engine

    def _sigma(self):
        s = self.likelihood.scale().reshape(-1)
        C_ = self._dims[0]
        return s.expand(C_) if s.numel() == 1 and C_ > 1 else s

    def _targets(self):
        y = self.y
        if self.mean is not None:
            y = y - self.mean(self.X).reshape(-1, 1)
        return y.reshape(-1)

    def _evaluate(self, packed, sigma, want_grad, targets=None):
        eng = self._eng()
        kind, p, dims = kernel_spec(self.kernel)
        if targets is None:
            targets = self._targets().detach()
        if self._rows is None or self._rows.owner is not eng:
            rows = eng.prepare(kind, {k: v.detach() for k, v in p.items()}, self.X, targets, self.data_variance)
            rows.owner = eng
            self._rows = rows
        elif self.mean is not None:
            self._rows.y = self._rows.sort_vec(targets, eng.device)
        out = eng.lml_grad_prepared(self._rows, packed.to(eng.device, torch.float64).contiguous(),
                                    sigma.to(eng.device, torch.float64).contiguous(), self.jitter, want_grad, check=False)
        info = int(out[1].item())            # the reference also synchronises here (float(loss))
        if info != 0:
            self._raise_cholesky(info, packed, sigma)
        self._factor_key = self._state_key(packed, sigma)
        return out

    def _state_key(self, packed, sigma):
        return (packed.detach().clone(), sigma.detach().clone())

    def _packed(self):
        kind, p, _ = kernel_spec(self.kernel)
        return torch.cat([p[n].reshape(-1) for n in _engine.PARAM_ORDER[kind]])

    def log_marginal_likelihood(self):
        """log p(y) (gpr/model.py:438-453); differentiable w.r.t. the raw parameters."""
        packed, sigma = self._packed(), self._sigma()
        targets = self._targets() if self.mean is not None else None
        if torch.is_grad_enabled() and (packed.requires_grad or sigma.requires_grad
                                        or (targets is not None and targets.requires_grad)):
            return _ExactLML.apply(packed, sigma, targets, self)
        return self._evaluate(packed.detach(), sigma.detach(), want_grad=False,
                              targets=None if targets is None else targets.detach())[0].clone()

    def _ensure_factor(self):
        """The reference re-factorises on every predict (gpr/model.py:463-469); here the factor of
        the last evaluation is reused when the parameters have not changed."""
        key = self._factor_key
        eng = self._eng()
        if key is not None and key[0] == "v" and eng._train is self._rows:
            groups = _param_tensors(self._kind, self.kernel) + [[self.likelihood.scale]]
            if key[1] == tuple((q.data_ptr(), q._version) for grp in groups for q in grp):
                return
            key = None
        packed, sigma = self._packed().detach(), self._sigma().detach()
        if self.mean is not None:
            key = None                       # alpha depends on the mean's parameters too: always re-evaluate
        if (key is None or eng._train is not self._rows or not torch.equal(key[0], packed)
                or not torch.equal(key[1], sigma)):
            self._evaluate(packed, sigma, want_grad=False)

    def K(self, X1, X2=None):
        with torch.inference_mode():
            return self.kernel(X1, X2)

    def predict_f(self, X, full=False):
        """Posterior mean / variance of f (gpr/model.py:455-483)."""
        with torch.no_grad():
            X = self._check_input(X)
            self._ensure_factor()
            mu, var = self._eng().predict(X, full=full)
            mu = mu.reshape(-1, 1)
            if self.mean is not None:
                mu = mu + self.mean(X).reshape(-1, 1)
            if not full:
                var = var.reshape(-1, 1)
            return mu, var

    def predict_y(self, X, ci=None, sigma=None, n=10000):       # gpr/model.py:322-344
        with torch.no_grad():
            X = self._check_input(X)
            mu, var = self.predict_f(X)
            if ci is None and sigma is not None:
                p = 0.5 * (1.0 + math.erf(sigma / math.sqrt(2.0)))
                ci = [1.0 - p, p]
            return self.likelihood.predict(X, mu, var, ci, sigma=sigma, n=n)

    def sample_f(self, Z, n=None, prior=False):                 # gpr/model.py:346-376
        with torch.no_grad():
            Z = self._check_input(Z)
            S = 1 if n is None else n
            if prior:
                mu = self.mean(Z).reshape(-1) if self.mean is not None else torch.zeros(Z.shape[0], device=Z.device, dtype=Z.dtype)
                var = self.kernel(Z)
            else:
                mu, var = self.predict_f(Z, full=True)
            var = var + self.jitter * var.diagonal().mean() * torch.eye(var.shape[0], device=var.device, dtype=var.dtype)
            samples = torch.distributions.multivariate_normal.MultivariateNormal(mu.reshape(-1), var).sample([S])
            return samples.squeeze() if n is None else samples

    def sample_y(self, Z, n=None):                              # gpr/model.py:378-401
        with torch.no_grad():
            Z = self._check_input(Z)
            S = 1 if n is None else n
            f = self.sample_f(Z, n=S)
            ys = self.likelihood.conditional_sample(Z, f.T).T
            return ys.squeeze() if n is None else ys
