"""Device-resident training loop for models built through ``B200Exact``.

The reference's loop (mogptk/model.py:563-565) is ``progress(i, float(gpr.loss())); optimizer.step()``: one host
synchronisation and ~10 small optimiser launches per iteration.  With the engine the step itself takes well
under a millisecond at BASELINE configs[1], so that host round trip is most of an iteration.  Here the whole
iteration -- parameter transforms, the fused exact-GP step (a replayed CUDA graph), the chain rule into
``p.grad`` and the Adam update of the raw leaves -- is enqueued by ONE C call for ``sync_every`` iterations at a
time (``mogp_train_adam``), and the host only synchronises between those chunks.

``losses`` / ``times`` / ``iters`` keep the reference's meaning (loss i is evaluated at the parameters before
update i, one extra evaluation after the last update, resumed runs append); ``times`` inside a chunk are
interpolated between the two synchronisations that bracket it.

Use either ``mogptk_b200.train.train(model, 'Adam', iters=..., lr=...)`` or ``mogptk_b200.install()``, which routes
``mogptk.Model.train`` here whenever the model was built with ``inference=B200Exact()`` and the request can run on
the device (Adam with lr / betas / eps only, no per-iteration ``error=`` callback); everything else falls through
to the reference's own ``train`` unchanged.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _cabi

_ADAM_KEYS = {"lr", "betas", "eps"}


def fused_adam_available(gpr_model, kwargs=None):
    """True when `gpr_model` is a mogptk_b200.gpr.Exact whose iteration can stay on the device."""
    from . import gpr as _gpr
    if not isinstance(gpr_model, _gpr.Exact):
        return False
    if kwargs is not None and not set(kwargs) <= _ADAM_KEYS:
        return False
    try:
        return gpr_model._fast_table() is not None
    except Exception:
        return False


def fit_adam(gpr_model, iters, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, sync_every=64, state=None, on_sync=None):
    """Run `iters` Adam iterations on the device.  Returns (losses[iters], times[iters]) as numpy arrays
    (times in seconds since the call started).  `state` (dict with exp_avg / exp_avg_sq / step) lets a caller
    continue an optimiser; by default a fresh one is used, as the reference creates one per train() call.
    Raises the model's CholeskyException at the first iteration whose covariance is not positive definite,
    with the parameters left exactly as they were at that iteration."""
    m = gpr_model
    fast = m._fast_table()
    if fast is None:
        raise NotImplementedError("this model needs torch autograd (mean function, priors or pegged parameters): "
                                  "use the reference's train()")
    entries, n_entries, P, plist = fast
    eng = m._eng()
    lib = eng.lib
    C_, Q, D = m._dims
    if m._rows is None or m._rows.owner is not eng:
        with torch.no_grad():
            m.log_marginal_likelihood()                  # parks x / y on the device (and validates the inputs)
    rows = m._rows
    T = P + C_
    dev = eng.device
    if state is None:
        state = {}
    if "exp_avg" not in state:
        state.update(exp_avg=torch.zeros(T, dtype=torch.float64, device=dev),
                     exp_avg_sq=torch.zeros(T, dtype=torch.float64, device=dev), step=0)
    # scratch lives in `state` so that a caller stepping one iteration at a time does not re-allocate (pinned memory is slow to get)
    sync_every = max(1, int(sync_every))
    chunk = min(sync_every, max(iters, 1))
    bufs = state.get("_bufs")
    if bufs is None or bufs["T"] != T or bufs["chunk"] < chunk or bufs["dev"] != dev:
        bufs = state["_bufs"] = {
            "T": T, "chunk": chunk, "dev": dev,
            "work": torch.empty(3 * (2 + T), dtype=torch.float64, device=dev),
            "losses": torch.empty(chunk, dtype=torch.float64, device=dev),
            "fail": torch.zeros(2, dtype=torch.int32, device=dev),
            "host_losses": torch.zeros(chunk, dtype=torch.float64).pin_memory(),
            "host_fail": torch.zeros(2, dtype=torch.int32).pin_memory()}
    work, losses_dev, fail = bufs["work"], bufs["losses"], bufs["fail"]
    host_losses, fail_host = bufs["host_losses"], bufs["host_fail"]
    fail.zero_()
    st = eng._stream()
    out_losses = np.empty(iters)
    times = np.zeros(iters)
    t0 = time.perf_counter()
    done, t_prev = 0, 0.0
    while done < iters:
        k = min(sync_every, iters - done)
        eng._check(lib.mogp_train_adam(
            eng.h, _cabi.KIND[m._kind], C_, Q, D, C.addressof(entries), n_entries, eng._p(rows.x), rows.off_p,
            eng._p(rows.y), eng._p(rows.dv), float(m.jitter), eng._p(work), eng._p(state["exp_avg"]),
            eng._p(state["exp_avg_sq"]), int(state["step"]), int(k), float(lr), float(betas[0]), float(betas[1]),
            float(eps), eng._p(losses_dev), eng._p(fail), st))
        host_losses[:k].copy_(losses_dev[:k], non_blocking=True)
        fail_host.copy_(fail, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()     # the chunk's one synchronisation
        t_now = time.perf_counter() - t0
        out_losses[done:done + k] = host_losses[:k].numpy()
        times[done:done + k] = t_prev + (t_now - t_prev) * (np.arange(1, k + 1) / k)
        t_prev = t_now
        eng._train, eng._kind = rows, m._kind
        m._factor_key = None                             # the leaves moved behind torch's version counters
        if int(fail_host[0]) != 0:
            # frozen at the failing iteration: re-evaluate there through the ordinary path, which raises
            m.loss()
            raise RuntimeError("mogp_train_adam reported info=%d at iteration %d but the re-evaluation succeeded"
                               % (int(fail_host[0]), done + int(fail_host[1])))
        state["step"] += k
        done += k
        if on_sync is not None:
            on_sync(done, out_losses[:done])
    return out_losses, times


def train(model, method="Adam", iters=500, verbose=False, error=None, plot=False, jit=None, sync_every=64, **kwargs):
    """Same contract as ``mogptk.Model.train`` (mogptk/model.py:440-606): returns (losses, errors) and maintains
    ``model.iters / times / losses / errors`` including the resume-and-append behaviour.  Runs on the device when
    it can (see module docstring) and otherwise calls the reference's own method."""
    ref_train = getattr(type(model), "_train_reference", None) or type(model).train
    fast = (str(method).lower() == "adam" and error is None and not plot
            and fused_adam_available(model.gpr, kwargs))
    if not fast:
        return ref_train(model, method=method, iters=iters, verbose=verbose, error=error, plot=plot, jit=jit, **kwargs)
    if verbose:
        print("Starting optimization using Adam (device-resident, %d iterations per synchronisation)" % sync_every)
        print("‣ Model: %s  ‣ Kernel: %s  ‣ Parameters: %d  ‣ Training points: %d  ‣ Iterations: %d"
              % (model.gpr.name(), model.gpr.kernel.name(), model.num_parameters(), model.num_training_points(), iters))
    t_start = time.time()
    had = model.times.shape[0]
    offset = had - 1 if had else 0
    new_losses, new_times = fit_adam(model.gpr, iters, sync_every=sync_every, **kwargs)
    final = float(model.gpr.loss())                      # the reference's closing progress(iters, self.loss())
    t_end = time.time() - t_start
    losses = np.concatenate([model.losses[:offset], new_losses, [final]])
    times = np.concatenate([model.times[:offset], new_times, [t_end]])
    errors = np.concatenate([model.errors[:offset] if model.errors.shape[0] >= offset else np.zeros(offset),
                             np.zeros(iters + 1)])
    model.iters = offset + iters
    model.times, model.losses = times, losses
    if verbose:
        print("  %d/%d  loss=%12g" % (model.iters, model.iters, final))
        print("Optimization finished in %.3f seconds" % t_end)
    return losses, errors


def install(mogptk=None):
    """Route the reference's exact-GP entry points that sit OUTSIDE the `inference=` seam to the engine (idempotent):

    * ``mogptk.Model.train`` -> :func:`train` (device-resident Adam loop when the request allows it);
    * ``mogptk.init.BNSE`` / the name ``BNSE`` imported into ``mogptk.data`` -> ``mogptk_b200.init.BNSE``;
    * ``Data.get_sm_estimation`` -> per-channel ``mogptk.SM`` built through ``B200Exact``;
    * ``DataSet.get_bnse_estimation`` -> the channels fitted concurrently.

    The reference's own functions stay reachable (``mogptk.Model._train_reference`` etc.) and are what runs for every
    request the engine does not cover; :func:`uninstall` restores them."""
    if mogptk is None:
        import mogptk
    from . import init as _init
    cls = mogptk.Model
    if getattr(cls, "_train_reference", None) is None:
        cls._train_reference = cls.train

        def _train(self, method="Adam", iters=500, verbose=False, error=None, plot=False, jit=None, **kwargs):
            return train(self, method=method, iters=iters, verbose=verbose, error=error, plot=plot, jit=jit, **kwargs)

        _train.__doc__ = cls._train_reference.__doc__
        cls.train = _train
    if getattr(mogptk, "_b200_saved", None) is None:
        import mogptk.data as mdata
        import mogptk.dataset as mdataset
        import mogptk.init as minit
        mogptk._b200_saved = {"init.BNSE": minit.BNSE, "data.BNSE": mdata.BNSE,
                              "Data.get_sm_estimation": mdata.Data.get_sm_estimation,
                              "DataSet.get_bnse_estimation": mdataset.DataSet.get_bnse_estimation}
        minit.BNSE = _init.BNSE
        mdata.BNSE = _init.BNSE

        def _get_sm_estimation(self, Q=1, method="LS", optimizer="Adam", iters=200, params={}):
            return _init.sm_estimation(self, Q, method, optimizer, iters, params)

        def _get_bnse_estimation(self, Q=1, n=1000, iters=200):
            return _init.bnse_estimation_concurrent(self, Q, n, iters)

        mdata.Data.get_sm_estimation = _get_sm_estimation
        mdataset.DataSet.get_bnse_estimation = _get_bnse_estimation
    return cls


def uninstall(mogptk=None):
    if mogptk is None:
        import mogptk
    cls = mogptk.Model
    if getattr(cls, "_train_reference", None) is not None:
        cls.train = cls._train_reference
        cls._train_reference = None
    saved = getattr(mogptk, "_b200_saved", None)
    if saved is not None:
        import mogptk.data as mdata
        import mogptk.dataset as mdataset
        import mogptk.init as minit
        minit.BNSE = saved["init.BNSE"]
        mdata.BNSE = saved["data.BNSE"]
        mdata.Data.get_sm_estimation = saved["Data.get_sm_estimation"]
        mdataset.DataSet.get_bnse_estimation = saved["DataSet.get_bnse_estimation"]
        mogptk._b200_saved = None
