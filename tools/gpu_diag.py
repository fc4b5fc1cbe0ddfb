#!/usr/bin/env python
"""Bring-up diagnostics for a gpurun call: each section runs in its own process (see
tools/gpu_diag.sh) so that a CUDA fault in one does not hide the others.
Usage: python tools/gpu_diag.py <section>"""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, float)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def sec_peak(eng):
    print("device:", torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).multi_processor_count, "SMs")
    for _ in range(2):
        d, f = eng.peak_fp64()
        print("peak fp64: DMMA %.2f TFLOP/s   DFMA %.2f TFLOP/s" % (d, f))
    import ctypes as C
    lat = (C.c_double * 16)()
    eng.lib.mogp_probe_latency(lat)
    names = ["DFMA", "DMUL", "rsqrt+add", "sqrt+add", "div+add", "LDS->addr", "DMMA", "SHFL64+add", "BAR(1 warp)"]
    print("dependent-chain latency (cycles/op): " + "  ".join("%s %.1f" % (n, lat[i]) for i, n in enumerate(names)))
    con = (C.c_double * 15)()
    eng.lib.mogp_probe_contention(con)
    print("DFMA chain latency with a DMMA stream on the same sub-partition (0/1/2/4 accumulators): " +
          "  ".join("chain %.1f dmma %.1f" % (con[3 * m], con[3 * m + 1]) for m in range(4)) +
          " | chain on warp 0, DMMA on warps 1,2,3,5,6,7: chain %.1f dmma %.1f (two warps share a sub-partition)" % (con[12], con[13]))
    iss = (C.c_double * 4)()
    eng.lib.mogp_probe_issue(iss)
    print("independent DFMA issue (8 chains per thread), cycles per instruction seen by warp 0 with 1/2/3/4 warps on its sub-partition: %.2f %.2f %.2f %.2f" % tuple(iss))
    print("8 dependent DMMAs after a scalar-fp64 gap of 0/64/256/1024 DFMAs: %.0f %.0f %.0f %.0f cycles" % (lat[10], lat[11], lat[12], lat[13]))


def sec_gemm(eng):
    for cfg in (3, 2):
        eng.lib.mogp_set_gemm_config(cfg)
        for ta in (0, 1):
            for tb in (0, 1):
                for (M, N, K) in [(64, 64, 16), (128, 128, 64), (192, 128, 256), (448, 384, 96)]:
                    g = torch.Generator().manual_seed(1)
                    A = torch.randn((K, M) if ta else (M, K), generator=g, dtype=torch.float64).cuda()
                    B = torch.randn((N, K) if tb else (K, N), generator=g, dtype=torch.float64).cuda()
                    C0 = torch.randn((M, N), generator=g, dtype=torch.float64).cuda()
                    ref = 0.7 * (A.T if ta else A) @ (B.T if tb else B) - 1.3 * C0
                    out = eng.dgemm(ta, tb, 0.7, A, B, -1.3, C0.clone())
                    torch.cuda.synchronize()
                    print("gemm cfg%d ta=%d tb=%d %4dx%4dx%4d relerr %.2e" % (cfg, ta, tb, M, N, K, rel(out, ref)))
        for n in (2048, 4096, 8192):
            A = torch.randn((n, n), dtype=torch.float64, device="cuda")
            B = torch.randn((n, n), dtype=torch.float64, device="cuda")
            Cm = torch.zeros((n, n), dtype=torch.float64, device="cuda")
            for ta, tb in ((0, 1), (0, 0), (1, 0)):
                med, mn = ev_time(lambda: eng.dgemm(ta, tb, 1.0, A, B, 0.0, Cm), reps=3, warm=1)
                print("gemm cfg%d ta=%d tb=%d n=%d: %.3f ms  %.2f TFLOP/s" % (cfg, ta, tb, n, mn, 2.0 * n ** 3 / mn / 1e9))
            med, mn = ev_time(lambda: torch.matmul(A, B, out=Cm), reps=3, warm=1)
            print("   cuBLAS dgemm n=%d: %.3f ms  %.2f TFLOP/s" % (n, mn, 2.0 * n ** 3 / mn / 1e9))
    eng.lib.mogp_set_gemm_config(0)


def spd(n, seed=0, shift=0.5):
    g = torch.Generator().manual_seed(seed)
    B = torch.randn((n, n + 8), generator=g, dtype=torch.float64)
    return B @ B.T / n + shift * torch.eye(n, dtype=torch.float64)


def sec_potrf(eng):
    print("panel variant (env):", os.environ.get("MOGP_PANEL_VARIANT", "default"), " trtri pipe:", os.environ.get("MOGP_TRTRI_PIPE", "default"))
    worst = 0.0
    for n in (64, 128, 200, 256, 640, 1024, 2048, 2176, 4096):
        A = spd(n, n)
        Lref = torch.linalg.cholesky(A)
        Ad = A.cuda().clone()
        info = eng.potrf_(Ad)
        L = torch.tril(Ad).cpu()
        e1, e2 = rel(L, Lref), rel(L @ L.T, A)
        worst = max(worst, e1, e2) if info == 0 else float("inf")
        print("potrf n=%4d info=%d  relerr(L) %.2e  relerr(LL^T) %.2e" % (n, info, e1, e2))
    A = spd(300, 5)
    A[150, 150] = -1.0
    bad = eng.potrf_(A.cuda().clone())
    print("potrf bad pivot -> info", bad, "(expect 151)")
    for n in (128, 384, 1024, 2048, 2176):
        A = spd(n, n + 1, 0.3)
        Ad = A.cuda().clone()
        Linv, Kinv, info = eng.trtri_kinv_(Ad)
        Lref = torch.linalg.cholesky(A)
        e = (rel(torch.tril(Ad).cpu(), Lref), rel(torch.tril(Linv).cpu(), torch.linalg.inv(Lref)),
             rel(torch.tril(Kinv).cpu(), torch.tril(torch.linalg.inv(A))))
        worst = max(worst, *e) if info == 0 else float("inf")
        print("trtri n=%4d info=%d relerr(L) %.2e relerr(Linv) %.2e relerr(Kinv) %.2e" % ((n, info) + e))
    print("POTRF_VERDICT %s worst %.2e" % ("OK" if worst < 1e-9 and bad == 151 else "FAIL", worst))
    for cfg in (0,):
        eng.lib.mogp_set_gemm_config(cfg)
        for n in (2048, 4096, 8192):
            A = spd(n, 1).cuda()
            W = A.clone()

            def run():
                W.copy_(A)
                eng.lib.mogp_potrf(eng.h, eng._p(W), n, n, None, eng._stream())
            def cp():
                W.copy_(A)
            t_all, _ = ev_time(run, reps=3, warm=1)
            t_cp, _ = ev_time(cp, reps=3, warm=1)
            t = t_all - t_cp
            print("potrf cfg%d n=%d: %.3f ms  %.2f TFLOP/s" % (cfg, n, t, n ** 3 / 3.0 / t / 1e9))
            if cfg == 0 and os.environ.get("MOGP_PANEL_VARIANT", "0") == "0":
                def cus():
                    torch.linalg.cholesky(A, out=W)
                t2, _ = ev_time(cus, reps=3, warm=1)
                print("   cuSOLVER potrf n=%d: %.3f ms  %.2f TFLOP/s" % (n, t2, n ** 3 / 3.0 / t2 / 1e9))
    eng.lib.mogp_set_gemm_config(0)


def sec_trtri(eng):
    for n in (128, 384, 1024, 2048):
        A = spd(n, n + 1, 0.3)
        Ad = A.cuda().clone()
        Linv, Kinv, info = eng.trtri_kinv_(Ad)
        Lref = torch.linalg.cholesky(A)
        print("trtri n=%4d info=%d relerr(L) %.2e relerr(Linv) %.2e relerr(Kinv) %.2e" % (
            n, info, rel(torch.tril(Ad).cpu(), Lref), rel(torch.tril(Linv).cpu(), torch.linalg.inv(Lref)),
            rel(torch.tril(Kinv).cpu(), torch.tril(torch.linalg.inv(A)))))


def sec_cov(eng):
    from conftest import golden_names, load_golden
    for name in golden_names():
        if name == "cfg3":
            continue
        try:
            g = load_golden(name)
            K = eng.K(g["kind"], g["params"], g["X"]).cpu().numpy()
            if "K_full" in g:
                e = rel(K, g["K_full"])
            else:
                idx = g["K_idx"]
                e = np.abs(K[idx[:, 0], idx[:, 1]] - g["K_val"]).max() / np.abs(g["K_val"]).max()
            kd = eng.K_diag(g["kind"], g["params"], g["X"]).cpu().numpy()
            Kfs = eng.K(g["kind"], g["params"], g["X"], g["Xs"]).cpu().numpy()
            e2 = np.abs(Kfs[::int(g["Kfs_row_stride"])] - g["Kfs_rows"]).max() / max(np.abs(g["K_diag"]).max(), 1e-300)
            print("K %-14s relerr %.2e  sym %s  kdiag==diag %s  cross relerr %.2e" % (
                name, e, np.array_equal(K, K.T), np.array_equal(kd, np.diagonal(K)), e2))
        except Exception:
            print("K %-14s FAILED" % name)
            traceback.print_exc()


def sec_lml(eng):
    from conftest import golden_names, load_golden
    for name in golden_names():
        try:
            g = load_golden(name)
            res = eng.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], True,
                               data_var=g.get("data_var"))
            errs = []
            for k, got in res["grad"].items():
                ref = g["gc_" + k]
                errs.append("%s %.1e" % (k[:4], np.abs(got.numpy().reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-12)))
            print("lml %-14s info=%d lml %.10f ref %.10f rel %.2e | grad %s" % (
                name, res["info"], res["lml"], float(g["lml"]), abs(res["lml"] - float(g["lml"])) / abs(float(g["lml"])),
                " ".join(errs)))
            if "pred_mu" in g:
                mu, var = eng.predict(g["Xs"])
                line = "    predict mu %.2e var %.2e" % (rel(mu, g["pred_mu"]), np.abs(var.cpu().numpy() - g["pred_var"]).max() / np.abs(g["pred_var"]).max())
                if "pred_cov" in g:
                    _, cov = eng.predict(g["Xs"], full=True)
                    line += " cov %.2e" % rel(cov, g["pred_cov"])
                print(line)
        except Exception:
            print("lml %-14s FAILED" % name)
            traceback.print_exc()


def sec_time(eng):
    from conftest import load_golden
    from mogptk_b200.engine import pack_params
    for cfg in (0,):
        eng.lib.mogp_set_gemm_config(cfg)
        for name in os.environ.get("DIAG_CFGS", "cfg1,cfg2,cfg4,cfg3").split(","):
            g = load_golden(name)
            rows = eng.prepare(g["kind"], g["params"], g["X"], g["y"])
            p = pack_params(g["kind"], g["params"], eng.device)
            sig = torch.tensor(g["sigma"], device=eng.device)
            N = g["X"].shape[0]
            eng.lib.mogp_set_graphs(0)
            tn, mn_ = ev_time(lambda: eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False), reps=9, warm=3)
            eng.lib.mogp_set_graphs(1)
            t1, m1 = ev_time(lambda: eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False), reps=9, warm=3)
            t0, m0 = ev_time(lambda: eng.lml_grad_prepared(rows, p, sig, g["jitter"], False, check=False), reps=9, warm=3)
            print("   graphs off: loss+grad %.3f ms (min %.3f)" % (tn, mn_))
            import ctypes as C
            eng.lib.mogp_set_profile(eng.h, 1)
            eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False)
            st = (C.c_float * 8)()
            ns = eng.lib.mogp_stage_times(eng.h, st)
            eng.lib.mogp_set_profile(eng.h, 0)
            print("   stages (sequential, profile mode): " + " ".join("%s %.3f" % (nm, st[i]) for i, nm in enumerate(
                ["kbuild", "potrf", "trtri", "solves", "kinv", "grad"][:ns])))
            lml_ref = float(g["lml"])
            r = eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=True)
            lml_got = float(r[0].item())
            print("   lml %.10f ref %.10f rel %.2e %s" % (lml_got, lml_ref, abs(lml_got - lml_ref) / abs(lml_ref),
                                                        "LML_OK" if abs(lml_got - lml_ref) <= 1e-8 * abs(lml_ref) else "LML_FAIL"))
            Kout = torch.empty((N, N), dtype=torch.float64, device=eng.device)
            def kb():
                eng.lib.mogp_kbuild(eng.h, {"MOSM": 0, "SM": 1, "CONV": 2}[g["kind"]], *rows.dims, eng._p(p), eng._p(rows.x),
                                    rows.off_p, None, None, None, None, 0.0, eng._p(Kout), N, eng._stream())
            tk, mk = ev_time(kb, reps=5, warm=2)
            print("time cfg%d %-5s N=%d: loss+grad %.3f ms (min %.3f) -> %.1f it/s | lml only %.3f ms | K full %.3f ms = %.0f GB/s" % (
                cfg, name, N, t1, m1, 1e3 / t1, t0, tk, 8.0 * N * N / mk / 1e6))
    eng.lib.mogp_set_gemm_config(0)


def sec_panel(eng):
    """Phase stamps (clock64) of the second panel step of an N=2048 factorisation, per panel variant.  The stamps and their
    stores perturb the kernel (the warp-specialised one more than the others): use `spans` for absolute kernel times."""
    import ctypes as C
    buf = (C.c_longlong * 64)()
    eng.lib.mogp_panel_debug(buf)          # arms the timestamps (second panel of the single-level sweep)
    n = 2048
    A = spd(n, 1).cuda()
    for variant in (2, 1, 0):
        eng.lib.mogp_set_panel_variant(variant)
        for _ in range(3):
            W = A.clone()
            eng.lib.mogp_potrf(eng.h, eng._p(W), n, n, None, eng._stream())
        torch.cuda.synchronize()
        eng.lib.mogp_panel_debug(buf)
        t = [int(v) for v in buf]
        if variant >= 1:
            print("ws%d panel n=%d: prologue %d | chain done stamps (delta): %s | tensor Xr-ready (rel. start): %s | total %d cycles" % (
                variant, n, t[1] - t[0], [t[2 + p] - (t[1] if p == 0 else t[1 + p]) for p in range(8)],
                [t[16 + p] - t[0] for p in range(8)], t[10] - t[0]))
            print("   chain u(p) = done(p) - Xr-ready(p): %s ; exchange = Xr-ready(p+1) - done(p): %s" % (
                [t[2 + p] - t[16 + p] for p in range(8)], [t[17 + p] - t[2 + p] for p in range(7)]))
            for base, pp in ((40, 0), (48, 3)):
                d = t[base:base + 6]
                print("   tensor warp 1 p=%d (start rel. kernel %d): prev-panel accumulation %d | published columns %d | barrier wait %d | tail+store %d | fence+arrive %d" % (
                    pp, d[0] - t[0], d[1] - d[0], d[2] - d[1], (d[3] - d[2]) if pp else 0, d[4] - (d[3] if pp else d[2]), d[5] - d[4]))
        else:
            print("panel n=%d: load %d | " % (n, t[1] - t[0]) + " ".join("p%d: f%d u%d" % (p, t[2 + 2 * p] - (t[1] if p == 0 else t[1 + 2 * p]), t[3 + 2 * p] - t[2 + 2 * p]) for p in range(8)) + " | store %d | total %d cycles" % (t[18] - t[17], t[18] - t[0]))
    eng.lib.mogp_set_panel_variant(int(os.environ.get("MOGP_PANEL_VARIANT", "2")))


def sec_exp(eng):
    """A/B of the Cholesky panel variants and the pipelined inverse in one process.
    EXP_COMBOS="variant:pipe,..." (default 0:0); prints a verdict line per combination."""
    import ctypes as C
    from conftest import load_golden
    from mogptk_b200.engine import pack_params
    combos = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("EXP_COMBOS", "0:0").split(",")]
    eng.lib.mogp_set_panel_pdl(int(os.environ.get("EXP_PDL", "1")))
    eng.lib.mogp_set_graph_max_np(int(os.environ.get("EXP_GRAPH_MAX_NP", "1000000")))
    eng.lib.mogp_set_two_level_above(int(os.environ.get("EXP_TWO_LEVEL_ABOVE", "2048")))
    print("pdl =", os.environ.get("EXP_PDL", "1"), "two-level above", os.environ.get("EXP_TWO_LEVEL_ABOVE", "2048"))
    names = os.environ.get("DIAG_CFGS", "cfg2,cfg4,cfg3").split(",")
    prepared = {}
    for name in names:
        g = load_golden(name)
        prepared[name] = (g, eng.prepare(g["kind"], g["params"], g["X"], g["y"]), pack_params(g["kind"], g["params"], eng.device),
                          torch.tensor(g["sigma"], device=eng.device))
    mats = {n: spd(n, n) for n in (128, 200, 640, 2048, 2176)}
    refs = {n: torch.linalg.cholesky(A) for n, A in mats.items()}
    invs = {n: torch.linalg.inv(refs[n]) for n in (128, 640, 2048, 2176)}
    for (v, pipe) in combos:
        eng.lib.mogp_set_panel_variant(v)
        eng.lib.mogp_set_trtri_pipe(pipe)
        tag = "v%d pipe%d" % (v, pipe)
        worst = 0.0
        for n, A in mats.items():
            Ad = A.cuda().clone()
            info = eng.potrf_(Ad)
            e1 = rel(torch.tril(Ad).cpu(), refs[n])
            worst = max(worst, e1) if info == 0 else float("inf")
            print("[%s] potrf n=%4d info=%d relerr(L) %.2e" % (tag, n, info, e1))
        Ab = spd(300, 5)
        Ab[150, 150] = -1.0
        bad = eng.potrf_(Ab.cuda().clone())
        for n in invs:
            Ad = mats[n].cuda().clone()
            Linv, Kinv, info = eng.trtri_kinv_(Ad)
            e = (rel(torch.tril(Ad).cpu(), refs[n]), rel(torch.tril(Linv).cpu(), invs[n]),
                 rel(torch.tril(Kinv).cpu(), torch.tril(invs[n].T @ invs[n])))
            worst = max(worst, *e) if info == 0 else float("inf")
            print("[%s] trtri n=%4d info=%d relerr(L) %.2e relerr(Linv) %.2e relerr(Kinv) %.2e" % ((tag, n, info) + e))
        print("[%s] VERDICT %s worst %.2e bad-pivot info %d (expect 151)" % (tag, "OK" if worst < 1e-9 and bad == 151 else "FAIL", worst, bad))
        for n in (2048, 4096, 8192):
            A = spd(n, 1).cuda()
            W = A.clone()

            def run():
                W.copy_(A)
                eng.lib.mogp_potrf(eng.h, eng._p(W), n, n, None, eng._stream())

            def cp():
                W.copy_(A)
            t_all, _ = ev_time(run, reps=5, warm=2)
            t_cp, _ = ev_time(cp, reps=5, warm=2)
            t = t_all - t_cp
            print("[%s] potrf n=%d: %.3f ms  %.2f TFLOP/s" % (tag, n, t, n ** 3 / 3.0 / t / 1e9))
        for name in names:
            g, rows, p, sig = prepared[name]
            N = g["X"].shape[0]
            t1, m1 = ev_time(lambda: eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False), reps=11, warm=4)
            eng.lib.mogp_set_profile(eng.h, 1)
            eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False)
            st = (C.c_float * 8)()
            ns = eng.lib.mogp_stage_times(eng.h, st)
            eng.lib.mogp_set_profile(eng.h, 0)
            r = eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False)
            lml_got, lml_ref = float(r[0].item()), float(g["lml"])
            gerr = 0.0
            P = p.numel()
            from mogptk_b200.engine import unpack_grads
            C_, Q, D = rows.dims
            gd = unpack_grads(g["kind"], C_, Q, D, r[2:2 + P].cpu())
            for k, got in gd.items():
                ref = g["gc_" + k]
                gerr = max(gerr, float(np.abs(got.numpy().reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-12)))
            ok = abs(lml_got - lml_ref) <= 1e-8 * abs(lml_ref) and gerr <= 1e-6 and int(r[1].item()) == 0
            print("[%s] step %-5s N=%d: %.3f ms (min %.3f) -> %.1f it/s | stages %s | lml rel %.1e grad %.1e %s" % (
                tag, name, N, t1, m1, 1e3 / t1, " ".join("%s %.3f" % (nm, st[i]) for i, nm in enumerate(
                    ["kbuild", "potrf", "trtri", "solves", "kinv", "grad"][:ns])),
                abs(lml_got - lml_ref) / abs(lml_ref), gerr, "STEP_OK" if ok else "STEP_FAIL"))
        sys.stdout.flush()
    eng.lib.mogp_set_panel_variant(2)
    eng.lib.mogp_set_trtri_pipe(1)
    eng.lib.mogp_set_panel_pdl(1)
    eng.lib.mogp_set_two_level_above(2048)


def sec_rowp(eng):
    """A/B of the row-wise pipelined inverse (ROWP_COMBOS="on:max_np:group,...") : numerics verdict + step time per config."""
    from conftest import load_golden
    from mogptk_b200.engine import pack_params, unpack_grads
    combos = [tuple(int(v) for v in c.split(":")) for c in
              os.environ.get("ROWP_COMBOS", "0:2048:1:4:1,1:2048:1:1:0,1:2048:1:4:1,1:2048:1:4:0,1:2048:1:8:1,1:2048:2:4:1,"
                                            "1:2048:2:8:1,1:4096:2:8:1,1:4096:4:16:1").split(",")]
    names = os.environ.get("DIAG_CFGS", "cfg1,mosm_mid,cfg2,cfg4").split(",")
    prepared = {}
    for name in names:
        g = load_golden(name)
        prepared[name] = (g, eng.prepare(g["kind"], g["params"], g["X"], g["y"]), pack_params(g["kind"], g["params"], eng.device),
                          torch.tensor(g["sigma"], device=eng.device))
    for combo in combos:
        assert eng.lib.mogp_set_rowpipe(*combo[:3]) == 0 and eng.lib.mogp_set_rowpipe_super(*combo[3:]) == 0
        tag = "rowp %d:%d:%d:%d:%d" % combo
        for n in (256, 640, 2048):
            A = spd(n, n)
            Ad = A.cuda().clone()
            Linv, Kinv, info = eng.trtri_kinv_(Ad)
            Li = torch.linalg.inv(torch.linalg.cholesky(A))
            print("[%s] trtri n=%4d info=%d relerr(Linv) %.2e relerr(Kinv) %.2e" % (
                tag, n, info, rel(torch.tril(Linv).cpu(), Li), rel(torch.tril(Kinv).cpu(), torch.tril(Li.T @ Li))))
        for name in names:
            g, rows, p, sig = prepared[name]
            N = g["X"].shape[0]
            t1, m1 = ev_time(lambda: eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False), reps=31, warm=5)
            r = eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False)
            lml_got, lml_ref = float(r[0].item()), float(g["lml"])
            P = p.numel()
            C_, Q, D = rows.dims
            gd = unpack_grads(g["kind"], C_, Q, D, r[2:2 + P].cpu())
            gerr = 0.0
            for k, got in gd.items():
                ref = g["gc_" + k]
                gerr = max(gerr, float(np.abs(got.numpy().reshape(ref.shape) - ref).max() / max(np.abs(ref).max(), 1e-12)))
            ok = abs(lml_got - lml_ref) <= 1e-8 * abs(lml_ref) and gerr <= 1e-6 and int(r[1].item()) == 0
            print("[%s] step %-8s N=%d: %.3f ms (min %.3f) -> %.1f it/s | lml rel %.1e grad %.1e %s" % (
                tag, name, N, t1, m1, 1e3 / t1, abs(lml_got - lml_ref) / abs(lml_ref), gerr, "STEP_OK" if ok else "STEP_FAIL"))
        sys.stdout.flush()
    eng.lib.mogp_set_rowpipe(1, 2048, 1)
    eng.lib.mogp_set_rowpipe_super(1, 0)


def sec_timeline(eng):
    """Global-timer timeline of one replayed step: stage stamps (mogp_set_stamps) + the span of every panel step, per
    row-pipeline setting (ROWP_COMBOS) and config (DIAG_CFGS)."""
    import ctypes as C
    from conftest import load_golden
    from mogptk_b200.engine import pack_params
    combos = [tuple(int(v) for v in c.split(":")) for c in os.environ.get("ROWP_COMBOS", "0:2048:1:4:1,1:2048:1:1:0,1:2048:1:4:1,1:2048:2:8:1").split(",")]
    names = os.environ.get("DIAG_CFGS", "cfg2").split(",")
    eng.lib.mogp_set_stamps(1)
    out = (C.c_ulonglong * 272)()
    for name in names:
        g = load_golden(name)
        rows = eng.prepare(g["kind"], g["params"], g["X"], g["y"])
        p = pack_params(g["kind"], g["params"], eng.device)
        sig = torch.tensor(g["sigma"], device=eng.device)
        nb = (g["X"].shape[0] + 127) // 128 * 2
        for combo in combos:
            eng.lib.mogp_set_rowpipe(*combo[:3])
            eng.lib.mogp_set_rowpipe_super(*combo[3:])
            for _ in range(6):
                eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False)
            torch.cuda.synchronize()
            eng.lib.mogp_panel_spans(1, out, 136)
            eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False)
            torch.cuda.synchronize()
            eng.lib.mogp_panel_spans(0, out, 136)
            t = [int(x) for x in out]
            t0 = t[256]
            st = [(t[256 + i] - t0) / 1e3 for i in range(6)]
            ent = [(t[2 * s] - t0) / 1e3 for s in range(nb)]
            ext = [(t[2 * s + 1] - t0) / 1e3 for s in range(nb)]
            spans = [ext[s] - ent[s] for s in range(nb)]
            gaps = [ent[s + 1] - ext[s] for s in range(nb - 1)]
            print("[timeline %s rowp %d:%d:%d:%d:%d] stamps us: kbuild done %.1f | chain %.1f -> %.1f (span mean %.2f, gap mean %.2f) | "
                  "potrf joined %.1f | solves done %.1f | kinv joined %.1f | end %.1f" % (
                      (name,) + combo + (st[1], ent[0], ext[nb - 1], float(np.mean(spans)), float(np.mean(gaps)), st[2], st[3], st[4], st[5])))
            print("    panel exit times: " + " ".join("%.0f" % x for x in ext))
            print("    spans: " + " ".join("%.1f" % x for x in spans))
            sys.stdout.flush()
    eng.lib.mogp_set_stamps(0)
    eng.lib.mogp_set_rowpipe(1, 2048, 1)
    eng.lib.mogp_set_rowpipe_super(1, 0)


def sec_rsizes(eng):
    """Recursive factor + inverse against the blocked sweep at sizes whose leaf is not 2048 (RS_SIZES = total rows): step time
    and the difference of LML / gradient between the two schedules."""
    from mogptk_b200 import synth
    from mogptk_b200.engine import pack_params
    sizes = [int(v) for v in os.environ.get("RS_SIZES", "4352,4608,5120,6144,7168").split(",")]
    for N in sizes:
        C_ = 4
        kind = "MOSM"
        X, y = synth.make_data(C_, [N // C_] * C_, seed=5)
        p, sigma = synth.make_params(kind, C_, 3, 1, seed=5)
        rows = eng.prepare(kind, p, X, y)
        pk = pack_params(kind, p, eng.device)
        sig = sigma.to(eng.device)
        res = {}
        for on in (1, 0):
            eng.lib.mogp_set_rchol(on, 4096, 2048)
            t, mn = ev_time(lambda: eng.lml_grad_prepared(rows, pk, sig, 1e-8, True, check=False), reps=7, warm=3)
            out = eng.lml_grad_prepared(rows, pk, sig, 1e-8, True, check=False).cpu()
            res[on] = (t, mn, out)
        a, b = res[1][2], res[0][2]
        gscale = float(b[2:].abs().max())
        eng.lib.mogp_set_rchol(1, 4096, 2048)
        print("rsizes N=%d leaf %d: recursive %.3f ms (min %.3f) | blocked %.3f ms (min %.3f) | info %d/%d lml rel diff %.1e grad diff %.1e" % (
            N, eng.lib.mogp_rchol_leaf_for((N + 127) // 128 * 128), res[1][0], res[1][1], res[0][0], res[0][1], int(a[1]), int(b[1]),
            abs(float(a[0] - b[0])) / abs(float(b[0])), float((a[2:] - b[2:]).abs().max()) / gscale))
        sys.stdout.flush()
    eng.lib.mogp_set_rchol(1, 4096, 2048)


def sec_gaps(eng):
    """Panel chain versus interference from the concurrent trailing updates (timing only: skip_bulk gives a wrong factor)."""
    for n in (2048, 4096):
        A = spd(n, 1).cuda()
        W = A.clone()

        def run():
            W.copy_(A)
            eng.lib.mogp_potrf(eng.h, eng._p(W), n, n, None, eng._stream())

        def cp():
            W.copy_(A)
        t_cp, _ = ev_time(cp, reps=5, warm=2)
        for v in (2, 1, 0):
            eng.lib.mogp_set_panel_variant(v)
            line = "gaps n=%d variant %d:" % (n, v)
            for skip in (0, 1):
                eng.lib.mogp_set_skip_bulk(skip)
                for cfg in (0, 2):
                    eng.lib.mogp_set_gemm_config(cfg)
                    t_all, _ = ev_time(run, reps=7, warm=2)
                    line += "  %s/%s %.3f ms" % ("no-bulk" if skip else "bulk", "32x64" if cfg == 0 else "64x64", t_all - t_cp)
            print(line)
    eng.lib.mogp_set_skip_bulk(0)
    eng.lib.mogp_set_gemm_config(0)
    eng.lib.mogp_set_panel_variant(2)


def sec_spans(eng):
    """Per panel step: kernel span seen from inside (global timer over all CTAs) and the gap to the next step."""
    import ctypes as C
    n = int(os.environ.get("SPAN_N", "2048"))
    nb = n // 64
    A = spd(n, 1).cuda()
    out = (C.c_ulonglong * (2 * nb))()
    for v, pdl in ((2, 0), (2, 2), (1, 0), (0, 0)):
        eng.lib.mogp_set_panel_variant(v)
        eng.lib.mogp_set_panel_pdl(pdl)
        print("--- variant %d pdl %d" % (v, pdl))
        for skip in (0, 1):
            eng.lib.mogp_set_skip_bulk(skip)
            for _ in range(3):
                W = A.clone()
                eng.lib.mogp_potrf(eng.h, eng._p(W), n, n, None, eng._stream())
            torch.cuda.synchronize()
            eng.lib.mogp_panel_spans(1, out, nb)
            W = A.clone()
            eng.lib.mogp_potrf(eng.h, eng._p(W), n, n, None, eng._stream())
            torch.cuda.synchronize()
            eng.lib.mogp_panel_spans(0, out, nb)
            t = [int(x) for x in out]
            spans = [(t[2 * s + 1] - t[2 * s]) / 1e3 for s in range(nb)]
            gaps = [(t[2 * s + 2] - t[2 * s + 1]) / 1e3 for s in range(nb - 1)]
            print("pdl requested %d active %d" % (pdl, eng.lib.mogp_get_panel_pdl()))
            print("spans n=%d variant %d %s: kernel span mean %.2f us (first 6: %s), gap to next step mean %.2f us (first 6: %s), chain total %.1f us" % (
                n, v, "no-bulk" if skip else "bulk", float(np.mean(spans)), " ".join("%.1f" % x for x in spans[:6]),
                float(np.mean(gaps)), " ".join("%.1f" % x for x in gaps[:6]), (t[2 * nb - 1] - t[0]) / 1e3))
    eng.lib.mogp_set_skip_bulk(0)
    eng.lib.mogp_set_panel_variant(2)


def sec_i8p(eng):
    """Product int8 tensor-pipe GEMM (csrc/i8mm.cu): error, time of slicing + MMA kernel, against the DMMA GEMM."""
    import ctypes as C
    out = (C.c_double * 4)()
    for ts in ([int(os.environ["I8_TS"])] if "I8_TS" in os.environ else [0, 1, 2]):
      eng.lib.mogp_set_i8_ts(1 if ts == 1 else 0)
      eng.lib.mogp_set_i8_wide(3 if ts == 2 else 0)
      print("--- %s" % ["128 x 64 tiles, one pass, both operands from shared memory (SS-form MMA)",
                        "128 x 64 tiles, A planes through tensor memory (tcgen05.cp + TS-form MMA)",
                        "128 x 128 tiles, two passes over the anti-diagonals (SS-form MMA)"][ts])
      for (M, N, K) in [(128, 128, 32), (256, 256, 256), (1024, 1024, 1024), (4096, 4096, 4096), (8192, 8192, 8192), (8192, 8192, 1024)]:
        for S in ((7,) if ts else (7, 8)):
            rc = eng.lib.mogp_i8_selftest(M, N, K, S, out)
            ops = 2.0 * M * N * K
            print("i8mm %5dx%5dx%5d S=%d rc=%d: rel. error %.2e | slicing+MMA %.3f ms, MMA kernel %.3f ms = %.1f TFLOP/s fp64-equivalent = %.0f TOP/s int8 | DMMA %.3f ms (%.1f TFLOP/s)" % (
                M, N, K, S, rc, out[0], out[1], out[2], ops / max(out[2], 1e-9) / 1e9, ops * S * (S + 1) / 2 / max(out[2], 1e-9) / 1e9,
                out[3], ops / max(out[3], 1e-9) / 1e9), flush=True)
    eng.lib.mogp_set_i8_ts(0)
    eng.lib.mogp_set_i8_wide(2)


def sec_gemm3(eng):
    """fp64 A B^T three ways per size: the DMMA kernel, the int8 tensor-pipe kernels (slicing included), cuBLAS dgemm."""
    import ctypes as C
    out = (C.c_double * 4)()
    print("%-22s %12s %12s %12s %12s   (TFLOP/s, fp64 or fp64-equivalent; int8 incl. operand slicing)" % (
        "M x N x K", "DMMA", "int8 128x64", "int8 128x128", "cuBLAS"))
    for (M, N, K) in [(1024, 1024, 1024), (2048, 2048, 2048), (4096, 4096, 4096), (8192, 8192, 8192), (8192, 8192, 1024),
                      (8192, 8192, 256), (4096, 4096, 256), (2048, 2048, 256)]:
        ops = 2.0 * M * N * K
        res = {}
        for name, wide in (("narrow", 0), ("wide", 3)):
            eng.lib.mogp_set_i8_wide(wide)
            eng.lib.mogp_i8_selftest(M, N, K, 7, out)
            res[name] = ops / max(out[1], 1e-9) / 1e9
            res["dmma"] = ops / max(out[3], 1e-9) / 1e9
        A = torch.randn((M, K), dtype=torch.float64, device="cuda")
        B = torch.randn((N, K), dtype=torch.float64, device="cuda")
        med, mn = ev_time(lambda: torch.matmul(A, B.T), reps=5, warm=2)
        print("%-22s %12.1f %12.1f %12.1f %12.1f" % ("%d x %d x %d" % (M, N, K), res["dmma"], res["narrow"], res["wide"], ops / mn / 1e9),
              flush=True)
        del A, B
    eng.lib.mogp_set_i8_wide(2)


def sec_gemmk(eng):
    """GEMM efficiency versus K and tile configuration (NT form, as in the Cholesky updates)."""
    shapes = [(8192, 8192, 64), (8192, 8192, 256), (8192, 8192, 1024), (4096, 4096, 256), (2048, 2048, 256),
              (2048, 2048, 64), (2048, 2048, 2048), (1024, 1024, 1024)]
    if os.environ.get("GEMMK_SHAPES"):
        shapes = [tuple(int(v) for v in t.split("x")) for t in os.environ["GEMMK_SHAPES"].split(",")]
    for (M, N, K) in shapes:
        A = torch.randn((M, K), dtype=torch.float64, device="cuda")
        B = torch.randn((N, K), dtype=torch.float64, device="cuda")
        Cm = torch.zeros((M, N), dtype=torch.float64, device="cuda")
        line = "gemm NT %5dx%5dx%5d beta=1:" % (M, N, K)
        if K == 64:                                   # the single-shot rank-64 kernel against the pipelined one (cfg columns)
            Cr = Cm.clone()
            eng.lib.mogp_set_gemm_k64(1)
            med, mn = ev_time(lambda: eng.dgemm(0, 1, -1.0, A, B, 1.0, Cm), reps=7, warm=3)
            eng.dgemm(0, 1, -1.0, A, B, 0.0, Cm)
            eng.lib.mogp_set_gemm_k64(0)
            eng.dgemm(0, 1, -1.0, A, B, 0.0, Cr)
            line += "  k64 %.3f ms %.1f TF (diff %.1e)" % (mn, 2.0 * M * N * K / mn / 1e9, float((Cm - Cr).abs().max()))
            Cm.zero_()
        for cfg in (3, 2, 4):
            eng.lib.mogp_set_gemm_config(cfg)
            med, mn = ev_time(lambda: eng.dgemm(0, 1, -1.0, A, B, 1.0, Cm), reps=5, warm=2)
            line += "  cfg%d %.3f ms %.1f TF" % (cfg, mn, 2.0 * M * N * K / mn / 1e9)
        med, mn = ev_time(lambda: torch.addmm(Cm, A, B.T, beta=1.0, alpha=-1.0, out=Cm), reps=5, warm=2)
        line += "  | cuBLAS %.3f ms %.1f TF" % (mn, 2.0 * M * N * K / mn / 1e9)
        print(line)
    eng.lib.mogp_set_gemm_config(0)
    eng.lib.mogp_set_gemm_k64(0)


def sec_thresh(eng):
    """Stage times of the exact-GP step versus the 32x64-tile threshold."""
    import ctypes as C
    from conftest import load_golden
    from mogptk_b200.engine import pack_params
    for name in ("cfg2", "cfg4", "cfg3"):
        g = load_golden(name)
        rows = eng.prepare(g["kind"], g["params"], g["X"], g["y"])
        p = pack_params(g["kind"], g["params"], eng.device)
        sig = torch.tensor(g["sigma"], device=eng.device)
        for thr in (0, 300, 700, 1400, 3000, 100000):
            eng.lib.mogp_set_small_tile_threshold(thr)
            t1, m1 = ev_time(lambda: eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False), reps=7, warm=2)
            eng.lib.mogp_set_profile(eng.h, 1)
            eng.lml_grad_prepared(rows, p, sig, g["jitter"], True, check=False)
            st = (C.c_float * 8)()
            ns = eng.lib.mogp_stage_times(eng.h, st)
            eng.lib.mogp_set_profile(eng.h, 0)
            print("thresh %-6s thr=%6d: step %.3f ms | potrf %.3f trtri %.3f kinv %.3f" % (name, thr, m1, st[1], st[2], st[4]))
    eng.lib.mogp_set_small_tile_threshold(1400)


def sec_train(eng):
    """Throughput of the reference-facing training loop: gpr.Exact.loss() + torch Adam step (what mogptk.Model.train runs)."""
    import time as _t
    from conftest import load_golden
    from test_host_layer import build_mirror
    from mogptk_b200 import gpr
    gpr.use_gpu(0)
    for name in ("cfg1", "cfg2", "cfg4"):
        g = load_golden(name)
        m, _ = build_mirror(g, eng, raw_from_golden=False)
        opt = torch.optim.Adam(m.parameters(), lr=0.01)
        for _ in range(5):
            float(m.loss()); opt.step()
        torch.cuda.synchronize()
        n = 60
        t0 = _t.perf_counter()
        for _ in range(n):
            l = float(m.loss())          # the reference's loop also synchronises on float(loss) (mogptk/model.py:384)
            opt.step()
        torch.cuda.synchronize()
        dt = (_t.perf_counter() - t0) / n
        t0 = _t.perf_counter()
        for _ in range(n):
            l = m.loss()
        torch.cuda.synchronize()
        dt2 = (_t.perf_counter() - t0) / n
        print("train %-5s: loss()+Adam %.3f ms/it (%.0f it/s) | loss() only, no per-it sync %.3f ms | final loss %.6f" % (
            name, dt * 1e3, 1 / dt, dt2 * 1e3, float(l)))


SECTIONS = {"rsizes": sec_rsizes, "timeline": sec_timeline, "rowp": sec_rowp, "i8p": sec_i8p, "gemm3": sec_gemm3, "spans": sec_spans, "gaps": sec_gaps, "exp": sec_exp, "train": sec_train, "thresh": sec_thresh, "gemmk": sec_gemmk, "panel": sec_panel, "peak": sec_peak, "gemm": sec_gemm, "potrf": sec_potrf, "trtri": sec_trtri, "cov": sec_cov,
            "lml": sec_lml, "time": sec_time}

if __name__ == "__main__":
    from mogptk_b200.engine import Engine
    name = sys.argv[1]
    t0 = time.time()
    eng = Engine(device=0, max_n=8192)
    print("=== section %s (engine up in %.1fs)" % (name, time.time() - t0), flush=True)
    try:
        SECTIONS[name](eng)
    except Exception:
        traceback.print_exc()
    torch.cuda.synchronize()
    print("=== section %s done in %.1fs" % (name, time.time() - t0), flush=True)
