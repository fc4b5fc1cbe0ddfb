#!/usr/bin/env python
"""bench.py -- log-marginal-likelihood iterations/sec (fp64) of the exact-GP hot path.

A "step" is one loss()-equivalent evaluation of one exact multi-output GP model:
K~ build + Cholesky + LML + full analytic gradient (mogptk/gpr/model.py:279-292, 438-453).
Workload at N GPUs = BASELINE.json configs[1] ("MOSM 4 channels, Q=5, N=2048, 1D, fp64")
per GPU; with N > 1 every rank evaluates an independent replica (another random restart,
seed = rank) and the ranks exchange only their final losses (one NCCL all-gather):
"replicas only", scaling = weak.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl b200|reference|reference-cuda]

Arms
  b200            this repository (default).  `value` = device-resident steps; `e2e` = the plug-in's Python API
                  (gpr.Exact.loss() + torch.optim.Adam.step() + float(loss), inputs copied host->device every step),
                  with the C-ABI host call, the device-resident training loop and the unmodified reference's own
                  train() loop over the plug-in reported beside it; `roofline` = the dominant stage of the step.
  reference       the UNMODIFIED reference (oracle/_ref, see oracle/build_ref.py): mogptk model -> gpr.Exact.loss()
                  (mogptk/gpr/model.py:279-292) on the host CPU, fixed thread policy (see REF_THREADS).  Falls back
                  to the oracle port only if the reference copy is missing (kind = "port").
  reference-cuda  the same unmodified reference with mogptk.gpr.use_gpu() (gpr/config.py:51-62): its own
                  PyTorch-CUDA path (ATen elementwise + cuSOLVER potrf) on the same B200.  Extra, clearly labelled.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mogptk_b200 import synth  # noqa: E402

METRIC = "lml_iters_per_sec"
UNIT = "it/s"
DMMA_PEAK_FALLBACK_TFLOPS = 37.15     # own probe (mogp_peak_fp64) on this pool, see DESIGN.md
REF_THREADS = 32                      # fixed CPU thread policy of the reference arm: min(32, host cores); round 1 probed
                                      # 8..all cores on this pool's hosts and torch-CPU was fastest at 16-32 threads
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_stage_traffic.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="cfg2", choices=sorted(synth.CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg3/cfg4 sub-records and the reference legs")
    return ap.parse_args()


def workload_name(cfg):
    kind, C, n, Q = synth.CONFIGS[cfg]
    return "%s %d channels, Q=%d, N=%d total (%d per channel), 1D input, Exact, fp64" % (kind, C, Q, C * n, n)


def config_dict(cfg):
    """Identical in every arm (the driver compares the dicts)."""
    return {"workload": workload_name(cfg), "config": cfg}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML every 10 ms; nvidia-smi fallback)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index=0):
        self.index, self.sm, self.mx, self.reasons = index, [], [], set()
        self._stop, self._t, self._nvml, self._h = threading.Event(), None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml, self._h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._nvml = None

    def _sample(self):
        if self._nvml is not None:
            n = self._nvml
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
            self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)))
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        else:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                  "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            v = [x.strip() for x in out.strip().split(",")]
            self.sm.append(float(v[0])); self.mx.append(float(v[1]))
            for name, x in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], v[2:6]):
                if x.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=3)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)), "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# --------------------------------------------------------------------------- the reference (CPU / its own CUDA path)
def reference_model(cfg, seed, device):
    """The unmodified reference's model for a BASELINE config: mogptk.MOSM/SM/CONV(dataset, Q) with the
    synthetic data and the hyper-parameters of mogptk_b200.synth assigned through Parameter.assign()."""
    from oracle import ref_loader
    mogptk = ref_loader.import_reference(device)
    kind, C, n, Q = synth.CONFIGS[cfg]
    _, p, sigma, X, y = synth.make_config(cfg, seed)
    ds = mogptk.DataSet()
    for c in range(C):
        msk = X[:, 0] == c
        ds.append(mogptk.Data(X[msk, 1], y[msk], name=str(c)))
    if kind == "MOSM":
        m = mogptk.MOSM(ds, Q=Q)
        k = m.gpr.kernel
        # (the constructor narrows `mean` to [1e-8, Nyquist] with a Sigmoid; re-assign inside those bounds)
        for name in ("weight", "mean", "variance", "delay", "phase"):
            getattr(k, name).assign(p[name])
    elif kind == "SM":
        m = mogptk.SM(ds, Q=Q)
        for c in range(C):
            for name in ("magnitude", "mean", "variance"):
                getattr(m.gpr.kernel[c], name).assign(p[name][c])
    else:
        m = mogptk.CONV(ds, Q=Q)
        for q in range(Q):
            for name in ("weight", "variance", "base_variance"):
                getattr(m.gpr.kernel[q], name).assign(p[name][q])
    m.gpr.likelihood.scale.assign(sigma)
    return mogptk, m


def time_reference(cfg, seed, steps, warmup, budget_s, device="cpu"):
    """(record, n_timed): median time of the reference's gpr.Exact.loss() (forward + autograd backward)."""
    from oracle import ref_loader
    cuda = device != "cpu"
    if not ref_loader.available():
        if cuda:
            return None, 0
        return cpu_port(cfg, seed, steps, warmup, budget_s)
    threads = min(REF_THREADS, os.cpu_count() or 1)
    torch.set_num_threads(threads)
    _, m = reference_model(cfg, seed, device)
    times, last = [], None
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        last = float(m.gpr.loss())                 # float() is the reference's own per-iteration sync (model.py:384)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    med = float(np.median(times))
    rec = {"value": 1.0 / med, "unit": UNIT, "cores": 0 if cuda else threads, "kind": "reference",
           "sample": "%d timed gpr.Exact.loss() calls (forward + autograd backward, mogptk/gpr/model.py:279-292) of the "
                     "unmodified reference (oracle/_ref) on %s, %s, median"
                     % (len(times), cfg, "its own PyTorch-CUDA path on this GPU (use_gpu)" if cuda else
                        "torch-CPU with %d threads (fixed policy: min(%d, host cores))" % (threads, REF_THREADS)),
           "ms_per_step": med * 1e3, "loss": last}
    return rec, len(times)


def cpu_port(cfg, seed, steps, warmup, budget_s):
    """Fallback when oracle/_ref is missing: the torch-CPU restatement in oracle/ (kind = "port")."""
    from oracle import mogp_oracle as orc
    kind, p, sigma, X, y = synth.make_config(cfg, seed)
    m = orc.RawModel(kind, p, sigma, X, y, 1e-8)
    threads = min(REF_THREADS, os.cpu_count() or 1)
    torch.set_num_threads(threads)
    times, last = [], None
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        last = float(m.loss().detach())
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    med = float(np.median(times))
    return {"value": 1.0 / med, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d timed loss() calls of the torch-CPU oracle port on %s (reference copy missing), median"
                      % (len(times), cfg), "ms_per_step": med * 1e3, "loss": last}, len(times)


def run_reference(args, device):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if device != "cpu" and not torch.cuda.is_available():
        emit({"impl": "reference-cuda", "unavailable": "no CUDA device"})
        return
    cb, n = time_reference(args.config, 0, args.steps, args.warmup, budget_s=150.0, device=device)
    if cb is None:
        emit({"impl": "reference-cuda", "unavailable": "oracle/_ref is missing"})
        return
    line = {"impl": "reference" if device == "cpu" else "reference-cuda", "metric": METRIC, "value": cb["value"],
            "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.config),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "loss": cb["loss"]}
    emit(line)


# --------------------------------------------------------------------------- B200 arm helpers
def mirror_model(cfg, seed, engine=None):
    """Stand-alone plug-in model (mogptk_b200.gpr mirror classes; no reference needed) for a config."""
    from mogptk_b200 import gpr
    kind, p, sigma, X, y = synth.make_config(cfg, seed)
    _, C, n, Q = synth.CONFIGS[cfg]
    if kind == "MOSM":
        k = gpr.MultiOutputSpectralMixtureKernel(Q=Q, output_dims=C, input_dims=1)
        for name in ("weight", "mean", "variance", "delay", "phase"):
            getattr(k, name).assign(p[name])
    elif kind == "SM":
        k = gpr.IndependentMultiOutputKernel([gpr.SpectralMixtureKernel(Q=Q, input_dims=1) for _ in range(C)], output_dims=C)
        for c in range(C):
            for name in ("magnitude", "mean", "variance"):
                getattr(k[c], name).assign(p[name][c])
    else:
        k = gpr.MixtureKernel(gpr.GaussianConvolutionProcessKernel(output_dims=C, input_dims=1), Q)
        for q in range(Q):
            for name in ("weight", "variance", "base_variance"):
                getattr(k[q], name).assign(p[name][q])
    m = gpr.Exact(k, X, y, variance=(sigma ** 2).tolist(), engine=engine)
    return m, X, y


def stage_times(eng, step, reps=3):
    import ctypes as C
    eng.lib.mogp_set_profile(eng.h, 1)
    for _ in range(reps):
        step()
    st = (C.c_float * 8)()
    ns = eng.lib.mogp_stage_times(eng.h, st)
    eng.lib.mogp_set_profile(eng.h, 0)
    names = ["kbuild", "potrf", "trtri", "solves", "kinv", "grad_finalize"]
    return {names[i]: round(float(st[i]), 4) for i in range(min(ns, len(names)))}


def int8_peak():
    """Dense int8 tensor peak to hold the tcgen05 kind::i8 stages against: MEASURED_PEAKS.json has no int8 figure, so twice
    its measured bf16 burst rate (the nominal ratio: 4.5 vs 2.25 PFLOP/s dense), else the nominal 4500 TOP/s."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return 2.0 * float(json.load(f)["bf16_tflops"]), "2 x measured bf16 burst (MEASURED_PEAKS.json)"
    except Exception:
        return 4500.0, "nominal dense int8"


def stage_rooflines(N, stages, dmma, hbm_gbs, traffic, i8=None, rchol_share=0.0):
    """Achieved rates of the stages against their rooflines (algorithmic work per SURVEY 8d).  i8 = (slices S, fraction of
    the triangular inverse's flops that runs on the int8 pipe) when the large GEMMs of this size take the tcgen05 path."""
    n3 = float(N) ** 3 / 3.0
    fused = stages.get("trtri", 1.0) < 0.02          # L^-1 pipelined behind the panel chain (N <= 4096)
    traffic = traffic or {}
    out = {}
    i8_peak, i8_src = int8_peak()

    def tensor(name, flops, ms, what, tkey=None, i8_frac=0.0):
        if ms and ms > 0:
            a = flops / (ms * 1e-3) / 1e12
            out[name] = {"bound": "tensor", "achieved": a, "peak": dmma, "unit": "TFLOP/s", "frac": a / dmma, "ms": ms,
                         "traffic": traffic.get(tkey or name), "what": what}
            if i8 and i8_frac > 0.0:
                S = i8[0]
                tops = i8_frac * flops * (S * (S + 1) / 2.0) / (ms * 1e-3) / 1e12
                out[name]["int8"] = {"achieved": tops, "peak": i8_peak, "unit": "TOP/s", "frac": tops / i8_peak,
                                     "what": "this stage runs (%.0f%% of its flops) on the int8 tensor pipe, tcgen05.mma kind::i8: "
                                             "every fp64 multiply-add is S (S + 1) / 2 = %d exact int8 multiply-adds (S = %d digit "
                                             "planes); achieved = those int8 operations / the whole stage time (operand slicing and "
                                             "the DMMA levels included); peak = %s.  `frac` above is the fp64-equivalent rate over "
                                             "the DMMA peak and may exceed 1." % (100 * i8_frac, S * (S + 1) // 2, S, i8_src)}
    if fused and rchol_share > 0.0:
        tensor("potrf_inverse", 2.0 * n3, stages.get("potrf"),
               "recursive Cholesky N^3/3 + L^-1 N^3/3: leaves by the blocked fp64-DMMA sweep with its row-wise pipelined inverse, every "
               "product above the leaves (%.0f%% of the flops) on the int8 tensor pipe" % (100 * rchol_share), "potrf_inverse",
               rchol_share)
    else:
        tensor("potrf_inverse" if fused else "potrf", (2.0 if fused else 1.0) * n3, stages.get("potrf"),
               "Cholesky N^3/3" + (" + L^-1 N^3/3 issued behind the panel chain (row-wise pipeline)" if fused else "") + " (fp64 DMMA)",
               "potrf_inverse")
    if not fused:
        tensor("trtri", n3, stages.get("trtri"), "L^-1 by level-batched block doubling, N^3/3", "trtri", i8[1] if i8 else 0.0)
    tensor("kinv", n3, stages.get("kinv"), "K^-1 = L^-T L^-1 (lower tiles), N^3/3, one launch", "kinv", 1.0 if i8 else 0.0)
    if stages.get("kbuild"):
        b = 8.0 * float(N) ** 2
        a = b / (stages["kbuild"] * 1e-3) / 1e9
        out["kbuild"] = {"bound": "hbm", "achieved": a, "peak": hbm_gbs, "unit": "GB/s", "frac": a / hbm_gbs,
                         "ms": stages["kbuild"], "traffic": traffic.get("kbuild"),
                         "what": "8 N^2 algorithmic bytes (SURVEY 8d convention; the fused step writes the lower half only) / "
                                 "stage time incl. the prep kernel; the kernel is fp64-pipe bound for Q >= 2 (DESIGN 4)"}
    return out


def i8_config(eng, N):
    """(S, int8 share of the triangular inverse's flops) when padded size N takes the int8 path, else None."""
    Np = (N + 127) // 128 * 128
    mn = int(eng.lib.mogp_get_i8_min_np())
    if mn <= 0 or Np < mn:
        return None
    tmin = int(eng.lib.mogp_get_i8_trtri_min())
    share, S_ = 0.0, 64
    while S_ < Np:                                   # level with block size S_ carries Np * S_^2 of the Np^3 / 3 flops
        if tmin > 0 and S_ >= tmin and Np % (2 * S_) == 0:
            share += float(Np) * S_ * S_
        S_ *= 2
    return int(eng.lib.mogp_get_i8_slices()), min(1.0, share / (float(Np) ** 3 / 3.0))


def rchol_share(eng, N):
    """Share of the factor + inverse flops above the leaves of the recursive scheme (0 when it does not apply)."""
    Np = (N + 127) // 128 * 128
    mn = int(eng.lib.mogp_get_i8_min_np())
    if mn <= 0 or Np < mn or not eng.lib.mogp_rchol_applies(Np):
        return 0.0
    leaf = C.c_longlong()
    eng.lib.mogp_get_rchol(None, C.byref(leaf))
    return 1.0 - (float(leaf.value) / Np) ** 2


def load_traffic(cfg):
    try:
        with open(TRAFFIC_FILE) as f:
            t = json.load(f)[cfg]["stages"]
        return {k: v["dram_bytes"] for k, v in t.items()}
    except Exception:
        return None


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def sub_record(cfg, device_index, dmma, hbm_gbs, steps=6):
    """A handful of device-resident steps of another BASELINE config on this GPU (north-star sizes, driver-run)."""
    from mogptk_b200.engine import Engine, pack_params
    kind, p, sigma, X, y = synth.make_config(cfg, 0)
    N = X.shape[0]
    eng = Engine(device=device_index, max_n=N)
    try:
        rows = eng.prepare(kind, p, X, y)
        packed = pack_params(kind, p, eng.device)
        sig = sigma.to(eng.device)
        P = packed.numel()

        def step(pk=packed):
            return eng.lml_grad_prepared(rows, pk, sig, 1e-8, True, check=False)
        for _ in range(3):
            out = step()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        pk = packed.clone()
        for e0, e1 in evs:
            e0.record()
            out = step(pk)
            e1.record()
            pk = pk - 1e-7 * out[2:2 + P]
        torch.cuda.synchronize()
        ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
        stages = stage_times(eng, step)
        # stand-alone Cholesky (mogp_potrf) of the same K~: the north-star "Cholesky TFLOP/s"
        K = eng.K(kind, p, X, sigma=sig, jitter=1e-8)
        A = K.clone()
        info = eng.potrf_(A)
        ts = []
        for _ in range(3):
            A.copy_(K)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            info_t = torch.zeros(1, dtype=torch.int32, device=eng.device)
            eng._check(eng.lib.mogp_potrf(eng.h, eng._p(A), N, A.stride(0), eng._p(info_t), eng._stream()))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        potrf_ms = float(min(ts))
        chol = (float(N) ** 3 / 3.0) / (potrf_ms * 1e-3) / 1e12
        return {"workload": workload_name(cfg), "steps": steps, "ms_per_step": ms, "value": 1e3 / ms, "unit": UNIT,
                "loss": -float(out[0]), "info": int(out[1]), "stage_ms": stages,
                "step": {"achieved": float(N) ** 3 / (ms * 1e-3) / 1e12, "frac": float(N) ** 3 / (ms * 1e-3) / 1e12 / dmma,
                         "unit": "TFLOP/s", "what": "N^3 algorithmic flop / median CUDA-event step time"},
                "cholesky": {"ms": potrf_ms, "achieved": chol, "peak": dmma, "frac": chol / dmma, "unit": "TFLOP/s", "info": info,
                             "what": "mogp_potrf alone on the same K~ (N^3/3 flop, best of 3, CUDA events)"},
                "stages": stage_rooflines(N, stages, dmma, hbm_gbs, load_traffic(cfg), i8_config(eng, N), rchol_share(eng, N))}
    finally:
        eng.close()


# --------------------------------------------------------------------------- B200 arm
def run_b200(args):
    from mogptk_b200 import replicas
    from mogptk_b200.engine import Engine, pack_params
    import mogptk_b200 as mb

    rank, world, local = replicas.env()
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the single JSON line (the banner goes to stdout)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    mb.gpr.use_gpu(local)
    replicas.init("nccl", torch.device("cuda", local))

    kind, p, sigma, X, y = synth.make_config(args.config, seed=rank)     # replica = another restart
    N = X.shape[0]
    eng = Engine(device=local, max_n=N)
    rows = eng.prepare(kind, p, X, y)
    dims = rows.dims
    P = int(sum(v.numel() for v in p.values()))
    packed = pack_params(kind, p, eng.device)
    sig = sigma.to(eng.device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)   # > 126 MB L2
    W = max(3, args.warmup)

    def step(pk):
        return eng.lml_grad_prepared(rows, pk, sig, 1e-8, True, check=False)

    def barrier():
        torch.cuda.synchronize()
        replicas.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    for _ in range(W):
        out = step(packed)
    torch.cuda.synchronize()
    info0 = int(out[1].item())
    launches0 = eng.lib.mogp_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pk = packed.clone()
    barrier()
    with ClockSampler(local) as clocks:
        t_wall0 = time.perf_counter()
        for e0, e1 in evs:
            flush.zero_()                                   # L2 flush, outside the per-step events
            e0.record()
            out = step(pk)
            e1.record()
            pk = pk - 1e-7 * out[2:2 + P]                   # move the parameters: nothing can be cached
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = eng.lib.mogp_launch_count() - launches0
    dev_ms = float(np.sum([a.elapsed_time(b) for a, b in evs]))
    final_loss = -float(out[0].item())
    dev_ms = replicas.max_over_ranks(dev_ms, eng.device)            # device time, max over ranks
    losses = replicas.all_gather_scalar(final_loss, eng.device)     # the path's only collective
    ms_per_step = dev_ms / args.steps
    value = world * 1e3 / ms_per_step

    stages = stage_times(eng, lambda: step(packed))

    # ---------------- e2e (headline): the plug-in's Python API.  Every step: x / y copied host -> device from pinned
    # memory, gpr.Exact.loss() (fills p.grad), torch.optim.Adam.step(), float(loss) (device -> host).
    model, _, _ = mirror_model(args.config, rank, engine=eng)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    with torch.no_grad():
        model.log_marginal_likelihood()                              # builds model._rows (x / y parked on the device)
    xh = torch.from_numpy(model._rows.x_host).pin_memory()
    yh = model._rows.y.cpu().pin_memory()

    def api_step():
        model._rows.x.copy_(xh, non_blocking=True)
        model._rows.y.copy_(yh, non_blocking=True)
        loss = model.loss()
        opt.step()
        return float(loss)
    for _ in range(W):
        api_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        api_loss = api_step()
    barrier()
    api_s = replicas.max_over_ranks(time.perf_counter() - t0, eng.device)
    e2e_val = world * args.steps / api_s
    h2d = 8 * (N * dims[2] + N)
    d2h = 16 + 8                                                      # [lml, info] read inside loss() + float(loss)

    # the same loop with torch's fused Adam (model.train(method='Adam', fused=True) passes it through, mogptk/model.py:556)
    opt_f = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)

    def api_step_fused():
        model._rows.x.copy_(xh, non_blocking=True)
        model._rows.y.copy_(yh, non_blocking=True)
        loss = model.loss()
        opt_f.step()
        return float(loss)
    for _ in range(W):
        api_step_fused()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        api_step_fused()
    barrier()
    e2e_fused_val = world * args.steps / replicas.max_over_ranks(time.perf_counter() - t0, eng.device)

    # the same per-step protocol (copies in, loss out, every step) with the optimiser update inside the C call:
    # Exact.fit_adam(1) = transforms + step + chain rule + Adam in one mogp_train_adam call, one synchronisation
    state = {}

    def api_step_device_adam():
        model._rows.x.copy_(xh, non_blocking=True)
        model._rows.y.copy_(yh, non_blocking=True)
        return float(model.fit_adam(1, lr=1e-3, sync_every=1, state=state)[0][0])
    for _ in range(W):
        api_step_device_adam()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        api_step_device_adam()
    barrier()
    e2e_dev_adam_val = world * args.steps / replicas.max_over_ranks(time.perf_counter() - t0, eng.device)

    # ---------------- beside it: the C-ABI host call (round-1 e2e), the device-resident training loop
    Pk = packed.cpu().numpy().copy()
    xhn, yhn = xh.numpy(), yh.numpy()
    ph = torch.from_numpy(Pk).pin_memory().numpy()
    sh = sigma.clone().pin_memory().numpy()
    oh = torch.empty(2 + P + dims[0], dtype=torch.float64).pin_memory().numpy()
    for _ in range(3):
        eng.lml_grad_host(kind, dims, ph, xhn, rows.chan_off, yhn, sh, 1e-8, True, oh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.lml_grad_host(kind, dims, ph, xhn, rows.chan_off, yhn, sh, 1e-8, True, oh)   # synchronises itself
        ph[:] = ph - 1e-7 * oh[2:2 + P]
    barrier()
    cabi_val = world * args.steps / replicas.max_over_ranks(time.perf_counter() - t0, eng.device)

    model2, _, _ = mirror_model(args.config, rank, engine=eng)
    mb.fit_adam(model2, 8, lr=1e-3, sync_every=8)
    barrier()
    t0 = time.perf_counter()
    fl, _ = mb.fit_adam(model2, args.steps, lr=1e-3, sync_every=64)
    barrier()
    fused_val = world * args.steps / replicas.max_over_ranks(time.perf_counter() - t0, eng.device)

    # ---------------- beside it: R independent replicas stacked on this GPU (restarts / sweeps; the panel chain of one
    # evaluation leaves two thirds of the SMs idle at this size), device-resident training loop, own handle + stream each
    conc = None
    if not args.no_extras and world == 1:
        try:
            conc = concurrent_replicas(args.config, local, 4, min(args.steps, 128))
        except Exception as e:
            conc = {"error": repr(e)}

    if rank == 0:
        try:
            dmma, dfma = eng.peak_fp64()
        except Exception:
            dmma, dfma = DMMA_PEAK_FALLBACK_TFLOPS, None
        hbm, hbm_src = hbm_peak()
        roofs = stage_rooflines(N, stages, dmma, hbm, load_traffic(args.config), i8_config(eng, N), rchol_share(eng, N))
        dom_name = max((k for k in roofs), key=lambda k: roofs[k]["ms"])
        dom = roofs[dom_name]
        ach = synth.flops_per_iteration(N) / (ms_per_step * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args.config),
            "details": {"parallelism": "replicas x%d" % world,
                        "l2": "flushed between timed steps (256 MiB memset outside the per-step CUDA events)",
                        "n_params": P + dims[0], "info": info0, "final_losses": losses,
                        "wall_s_incl_flush": round(t_wall, 4)},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "plug-in Python API per step: x, y host->device from pinned memory, gpr.Exact.loss() (raw-space "
                            "p.grad filled), torch.optim.Adam.step(), float(loss); wall clock, max over ranks.  loss() returns "
                            "when the step has published [lml, info] (mapped pinned memory, right after the solves; a Cholesky "
                            "failure still raises there); K^-1, the gradient reduction and the chain rule finish in stream "
                            "order under the optimiser's host-side work (MOGP_EARLY_LOSS=0: synchronise instead)",
                    "last_loss": api_loss,
                    "fused_optimizer": {"value": e2e_fused_val, "unit": UNIT,
                                        "what": "the same per-step loop with torch.optim.Adam(fused=True) (one optimiser kernel "
                                                "instead of ~10), the option mogptk.Model.train(method='Adam', fused=True) passes on"},
                    "per_step_device_adam": {"value": e2e_dev_adam_val, "unit": UNIT,
                                             "what": "same per-step copies and loss read-back, but loss + Adam update in one call: "
                                                     "gpr.Exact.fit_adam(1) (mogp_train_adam, one synchronisation per step)"},
                    "c_abi_host_call": {"value": cabi_val, "unit": UNIT, "h2d_bytes_per_step": 8 * (P + dims[0] + N * dims[2] + N),
                                        "d2h_bytes_per_step": 8 * (2 + P + dims[0]),
                                        "what": "mogp_lml_grad_host: params, sigma, x, y host->device, step, LML + gradient "
                                                "device->host, synchronised (round 1's e2e)"},
                    "device_resident_training": {"value": fused_val, "unit": UNIT, "sync_every": 64, "last_loss": float(fl[-1]),
                                                 "what": "mogptk_b200.fit_adam / mogp_train_adam: transforms + step + chain rule + "
                                                         "Adam update enqueued by one C call per 64 iterations, one host "
                                                         "synchronisation per chunk (what mogptk_b200.install() routes "
                                                         "mogptk.Model.train('Adam') to); data resident, no per-step copies"}},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": {
                "kernel": dom_name, "bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"],
                "unit": dom["unit"], "frac": dom["frac"], "traffic": dom["traffic"],
                "what": "dominant stage of the benched step by CUDA-event time inside the library (%s: %.4f of %.4f ms "
                        "profiled sequentially): %s. achieved = algorithmic flop of the stage / its duration; peak = fp64 "
                        "tensor-pipe (DMMA m8n8k4) probe measured in this run (MEASURED_PEAKS.json has no fp64 figure; "
                        "DFMA probe %.2f TFLOP/s); hbm peak %s. traffic = dram read+write bytes of the stage's launches from "
                        "the committed ncu pass (profiles/r02_stage_traffic.json), null if that file is absent."
                        % (dom_name, dom["ms"], sum(stages.values()), dom["what"], dfma or 0.0, hbm_src),
                "stages": roofs,
                "step": {"achieved": ach, "frac": ach / dmma,
                         "what": "whole step: N^3 algorithmic flop (potrf N^3/3 + inverse 2N^3/3) / CUDA-event step time"},
                "stage_ms": stages},
        }
        if conc is not None:
            line["e2e"]["concurrent_replicas"] = conc
        if not args.no_extras and world == 1:
            extra = {}
            for cfg in ("cfg3", "cfg4"):
                if cfg == args.config:
                    continue
                try:
                    extra[cfg] = sub_record(cfg, local, dmma, hbm)
                except Exception as e:                      # never lose the headline line to a sub-record
                    extra[cfg] = {"error": repr(e)}
            line["configs"] = extra
            # the unmodified reference driving the plug-in through its own train() loop, and its own CUDA path
            try:
                line["reference_seam"] = reference_seam(args.config, min(args.steps, 200))
            except Exception as e:
                line["reference_seam"] = {"error": repr(e)}
            try:
                rc, _ = time_reference(args.config, 0, 10, 2, budget_s=60.0, device="cuda:%d" % local)
                line["reference_cuda"] = rc if rc is not None else {"unavailable": "oracle/_ref is missing"}
            except Exception as e:
                line["reference_cuda"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world == 1:
            cb, _ = time_reference(args.config, 0, 20, 2, budget_s=25.0, device="cpu")
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        emit(line)
    replicas.finish()


def concurrent_replicas(cfg, device_index, R, iters):
    """Aggregate it/s of R independent models (seeds 0..R-1) trained concurrently on one GPU."""
    from mogptk_b200 import replicas
    from mogptk_b200.engine import Engine
    N = synth.make_config(cfg, 0)[3].shape[0]
    engines = [Engine(device=device_index, max_n=N) for _ in range(R)]
    try:
        models = [mirror_model(cfg, r, engine=engines[r])[0] for r in range(R)]
        replicas.train_restarts(models, 8, lr=1e-3, sync_every=8)             # warm-up: graph capture per handle
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hist = replicas.train_restarts(models, iters, lr=1e-3, sync_every=64)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        one = [Engine(device=device_index, max_n=N)]
        try:
            m1 = mirror_model(cfg, 0, engine=one[0])[0]
            replicas.train_restarts([m1], 8, lr=1e-3, sync_every=8)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            replicas.train_restarts([m1], iters, lr=1e-3, sync_every=64)
            torch.cuda.synchronize()
            dt1 = time.perf_counter() - t0
        finally:
            one[0].close()
        return {"replicas": R, "iters_each": iters, "value": R * iters / dt, "single_replica_value": iters / dt1, "unit": UNIT,
                "final_losses": [float(h[-1]) for h in hist],
                "what": "R independent models (seeds 0..R-1: restarts) on ONE GPU, each with its own workspace handle, CUDA "
                        "stream and host thread, device-resident Adam loop (64 iterations per synchronisation); aggregate "
                        "iterations/s over all replicas, wall clock; single_replica_value = the same loop with R = 1"}
    finally:
        for e in engines:
            e.close()


def reference_seam(cfg, iters):
    """mogptk.MOSM(dataset, Q, inference=B200Exact()).train('Adam') -- the UNMODIFIED reference's loop
    (mogptk/model.py:563-566) over the plug-in, and the same call after mogptk_b200.install()."""
    from oracle import ref_loader
    import mogptk_b200 as mb
    if not ref_loader.available():
        return {"unavailable": "oracle/_ref is missing"}
    mogptk = ref_loader.import_reference("cuda")
    kind, C, n, Q = synth.CONFIGS[cfg]
    _, p, sigma, X, y = synth.make_config(cfg, 0)

    def build():
        ds = mogptk.DataSet()
        for c in range(C):
            msk = X[:, 0] == c
            ds.append(mogptk.Data(X[msk, 1], y[msk], name=str(c)))
        m = getattr(mogptk, kind)(ds, Q=Q, inference=mb.B200Exact())
        if kind == "MOSM":
            for name in ("weight", "mean", "variance"):
                getattr(m.gpr.kernel, name).assign(p[name])
        return m
    m = build()
    m.train(method="Adam", iters=5, lr=1e-3, verbose=False, jit=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    m.train(method="Adam", iters=iters, lr=1e-3, verbose=False, jit=False)
    torch.cuda.synchronize()
    plain = iters / (time.perf_counter() - t0)
    m2 = build()
    mb.install(mogptk)
    try:
        m2.train(method="Adam", iters=5, lr=1e-3, verbose=False, jit=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m2.train(method="Adam", iters=iters, lr=1e-3, verbose=False, jit=False)
        torch.cuda.synchronize()
        inst = iters / (time.perf_counter() - t0)
    finally:
        mb.uninstall(mogptk)
    return {"unit": UNIT, "iters": iters,
            "model_train_adam": {"value": plain, "what": "mogptk.Model.train('Adam') of the unmodified reference driving "
                                 "B200Exact: float(gpr.loss()) + torch.optim.Adam.step() per iteration (model.py:563-565)"},
            "model_train_adam_installed": {"value": inst, "what": "the same call after mogptk_b200.install(): device-resident "
                                           "loop, 64 iterations per synchronisation"},
            "final_loss": float(m.losses[-1]), "final_loss_installed": float(m2.losses[-1])}


class _StdoutToStderr:
    """Everything the libraries print while the bench runs (e.g. the NCCL version banner at N > 1, which goes to file
    descriptor 1) is sent to stderr: stdout carries the ONE JSON line, printed after the descriptor is restored."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None

    def __exit__(self, *exc):
        self.restore()


_REDIRECT = None


def emit(line):
    """The single JSON line on the real stdout."""
    if _REDIRECT is not None:
        _REDIRECT.restore()
    sys.stdout.write(json.dumps(line) + "\n")
    sys.stdout.flush()


if __name__ == "__main__":
    a = parse()
    with _StdoutToStderr() as _REDIRECT:
        if a.impl == "reference":
            run_reference(a, "cpu")
        elif a.impl == "reference-cuda":
            run_reference(a, "cuda")
        else:
            run_b200(a)
