"""GPU tests of the scheduling knobs of the Cholesky / inverse stage: every panel-step variant, the pipelined and
the level-batched triangular inverse, programmatic dependent launch on and off, graph replay on and off must all
reproduce the reference (same tolerances as test_gpu_parity.py: LML rtol 1e-8, gradients <= 1e-6 of the largest
entry).  The defaults are what the rest of the suite runs; these keep the alternatives -- which are also the
fallbacks when a launch attribute is not supported -- honest."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture()
def knobs(engine):
    lib = engine.lib

    def set_(variant=2, pipe=1, pdl=1, graphs=1, rowpipe=(1, 4096, 1, 1, 0), kinv=(0, 4)):
        lib.mogp_set_panel_variant(variant)
        lib.mogp_set_trtri_pipe(pipe)
        lib.mogp_set_panel_pdl(pdl)
        lib.mogp_set_graphs(graphs)
        assert lib.mogp_set_rowpipe(*rowpipe[:3]) == 0 and lib.mogp_set_rowpipe_super(*rowpipe[3:]) == 0
        assert lib.mogp_set_rowpipe_kinv(kinv[0]) == 0 and lib.mogp_set_rowpipe_wmin(kinv[1]) == 0
    yield set_
    set_()          # back to the defaults for the rest of the session


def _check(engine, g, reps=3):
    # three evaluations: plain run, graph capture, graph replay
    for _ in range(reps):
        res = engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], True,
                              data_var=g.get("data_var"))
    assert res["info"] == 0
    assert abs(res["lml"] - float(g["lml"])) <= 1e-8 * abs(float(g["lml"]))
    for k, got in res["grad"].items():
        ref = g["gc_" + k]
        scale = max(float(np.abs(ref).max()), 1e-12)
        assert float(np.abs(got.numpy().reshape(ref.shape) - ref).max()) <= 1e-6 * scale, k


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("pipe", [0, 1])
@pytest.mark.parametrize("name", ["mosm_mid", "cfg2"])
def test_panel_variants_and_inverse_schedules(engine, knobs, name, variant, pipe):
    knobs(variant=variant, pipe=pipe)
    _check(engine, load_golden(name))


@pytest.mark.parametrize("rowpipe", [(0, 2048, 2, 4, 1), (1, 2048, 1, 1, 0), (1, 2048, 1, 4, 1), (1, 2048, 2, 8, 0), (1, 2048, 4, 4, 1),
                                     (1, 4096, 8, 16, 1), (1, 4096, 2, 6, 1)])
@pytest.mark.parametrize("kinv", [(0, 4), (1, 4), (2, 4), (2, 1)])
@pytest.mark.parametrize("name", ["mosm_small", "mosm_mid", "cfg1", "cfg2", "cfg4"])
def test_rowwise_pipelined_inverse(engine, knobs, name, rowpipe, kinv):
    """The row-wise pipeline (Linv by row groups and K^-1 by rank updates behind the panel chain; on, max_np, group, super-group,
    taper) against the block-doubling pipeline (rowpipe off): group / super-group sizes, ragged last groups (mosm_*) and the
    two-level Cholesky sweep (cfg4 with max_np 4096); K^-1 afterwards (kinv mode 0), accumulated per super-group (1) or in
    halving chunks (2, smallest chunk 4 / 1 blocks)."""
    knobs(rowpipe=rowpipe, kinv=kinv)
    _check(engine, load_golden(name))


@pytest.mark.parametrize("group,sup,taper", [(1, 1, 0), (1, 4, 1), (2, 4, 0), (2, 8, 1), (4, 8, 1), (8, 8, 0)])
@pytest.mark.parametrize("kinv", [(1, 4), (2, 4), (2, 2)])
@pytest.mark.parametrize("n", [256, 640, 1152, 2048])
def test_rowwise_trtri_kinv(engine, knobs, n, group, sup, taper, kinv):
    knobs(rowpipe=(1, 2048, group, sup, taper), kinv=kinv)
    gen = torch.Generator().manual_seed(n + group + sup)
    B = torch.randn((n, n + 8), generator=gen, dtype=torch.float64)
    A = B @ B.T / n + 0.3 * torch.eye(n, dtype=torch.float64)
    for _ in range(2):            # the second run finds stale values in the scratch / accumulation buffers
        Linv, Kinv, info = engine.trtri_kinv_(A.cuda().clone())
        assert info == 0
        Linv_ref = torch.linalg.inv(torch.linalg.cholesky(A))
        assert float((torch.tril(Linv).cpu() - Linv_ref).abs().max() / Linv_ref.abs().max()) < 1e-10
        Kinv_ref = torch.tril(torch.linalg.inv(A))
        assert float((torch.tril(Kinv).cpu() - Kinv_ref).abs().max() / Kinv_ref.abs().max()) < 1e-10


@pytest.mark.parametrize("pdl,graphs", [(0, 1), (2, 1), (2, 0), (1, 0)])
@pytest.mark.parametrize("name", ["cfg2", "cfg4"])
def test_dependent_launch_and_graph_modes(engine, knobs, name, pdl, graphs):
    knobs(pdl=pdl, graphs=graphs)
    _check(engine, load_golden(name))


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("n", [64, 200, 1024, 4224, 4736])
def test_potrf_variants_against_lapack(engine, knobs, variant, n):
    """4224 / 4736 rows take the two-level sweep; at 4736 the early panel steps use 64 own rows per CTA and the late ones 32."""
    knobs(variant=variant)
    gen = torch.Generator().manual_seed(n)
    B = torch.randn((n, n + 8), generator=gen, dtype=torch.float64)
    A = B @ B.T / n + 0.5 * torch.eye(n, dtype=torch.float64)
    Ad = A.cuda().clone()
    assert engine.potrf_(Ad) == 0
    L = torch.tril(Ad).cpu()
    Lref = torch.linalg.cholesky(A)
    assert float((L - Lref).abs().max() / Lref.abs().max()) < 1e-12


def test_failed_graph_capture_falls_back_to_an_eager_run(engine, knobs):
    """If the step cannot be captured (e.g. no programmatic-dependent-launch edges in graphs on an older driver) the
    same call must still return the right numbers, and later calls run eagerly."""
    knobs()
    lib = engine.lib
    g = load_golden("mosm_mid")
    lib.mogp_test_fail_capture(1)
    try:
        _check(engine, g, reps=4)            # plain run, failed capture -> eager, eager, eager
    finally:
        lib.mogp_test_fail_capture(0)
        knobs()                              # re-enables graphs and dependent launches
    _check(engine, g, reps=3)                # and the replayed graph still agrees afterwards


# ------------------------------------------------------------------ fp64 GEMM on the int8 tensor pipe (tcgen05.mma kind::i8)
@pytest.mark.parametrize("ts", [0, 1, 2])
@pytest.mark.parametrize("M,N,K,S,tol", [(128, 128, 32, 7, 1e-12), (256, 384, 4096, 7, 1e-12), (1024, 1024, 1024, 8, 1e-13),
                                          (2048, 2048, 2048, 7, 1e-12)])
def test_int8_tensor_pipe_gemm_against_dmma(engine, M, N, K, S, tol, ts):
    """A B^T from S signed 7-bit digit planes per operand (Ozaki slicing, exact int32 accumulation in TMEM) against the
    fp64 DMMA GEMM: relative to the largest entry the difference is ~1e-14 at S = 7 and ~fp64 rounding at S = 8."""
    import ctypes as C
    out = (C.c_double * 4)()
    default_ts, default_wide = engine.lib.mogp_get_i8_ts(), engine.lib.mogp_get_i8_wide()
    try:
        # 0: 128 x 64 tiles, both operands from shared memory; 1: A planes through tensor memory; 2: 128 x 128 tiles, two passes
        engine.lib.mogp_set_i8_ts(1 if ts == 1 else 0)
        engine.lib.mogp_set_i8_wide(3 if ts == 2 else 0)
        rc = engine.lib.mogp_i8_selftest(M, N, K, S, out)
    finally:
        engine.lib.mogp_set_i8_ts(default_ts)
        engine.lib.mogp_set_i8_wide(default_wide)
    assert rc == 0
    assert 0.0 <= out[0] < tol, out[0]


def test_int8_kinv_equals_dmma_kinv(engine):
    """K^-1 = L^-T L^-1 of an SPD matrix at n = 4096 through the int8 path (default for n >= 4096) and through DMMA."""
    import torch
    n = 4096
    torch.manual_seed(3)
    B = torch.randn(n, n, dtype=torch.float64, device=engine.device)
    K = B @ B.T / n + torch.eye(n, dtype=torch.float64, device=engine.device)
    try:
        engine.lib.mogp_set_i8(0, 7)
        _, Kd, info_d = engine.trtri_kinv_(K.clone())
        engine.lib.mogp_set_i8(2048, 7)
        _, K7, info_7 = engine.trtri_kinv_(K.clone())
        engine.lib.mogp_set_i8(4096, 8)
        _, K8, info_8 = engine.trtri_kinv_(K.clone())
    finally:
        engine.lib.mogp_set_i8(2048, 7)
    assert info_d == 0 and info_7 == 0 and info_8 == 0
    ref = torch.linalg.inv(K)
    scale = float(ref.abs().max())
    tril = torch.tril(torch.ones(n, n, dtype=torch.bool, device=engine.device))
    for got, tol in ((Kd, 1e-11), (K7, 1e-11), (K8, 1e-11)):
        assert float((got - ref)[tril].abs().max()) <= tol * scale
    assert float((K7 - Kd)[tril].abs().max()) <= 1e-12 * scale
    assert float((K8 - Kd)[tril].abs().max()) <= 1e-13 * scale


@pytest.mark.parametrize("n", [4096, 8192])
def test_int8_triangular_inverse_levels_equal_dmma(engine, n):
    """L^-1 by block doubling with the large levels (block size >= 1024) on the int8 tensor pipe against the all-DMMA
    recursion; n = 4096 with the pipelined inverse switched off so that the level-batched path runs."""
    import torch
    torch.manual_seed(5)
    B = torch.randn(n, n // 2, dtype=torch.float64, device=engine.device)
    K = B @ B.T / n + 0.5 * torch.eye(n, dtype=torch.float64, device=engine.device)
    del B
    lib = engine.lib
    try:
        lib.mogp_set_trtri_pipe(0)
        lib.mogp_set_i8(0, 7)
        Ld, Kd, info_d = engine.trtri_kinv_(K.clone())
        lib.mogp_set_i8(2048, 7)
        L7, K7, info_7 = engine.trtri_kinv_(K.clone())
    finally:
        lib.mogp_set_i8(2048, 7)
        lib.mogp_set_trtri_pipe(1)
    assert info_d == 0 and info_7 == 0
    tril = torch.tril(torch.ones(n, n, dtype=torch.bool, device=engine.device))
    sl, sk = float(Ld.abs().max()), float(Kd[tril].abs().max())
    # S = 7 digit planes carry ~3e-14 of the largest entry per product; two int8 levels and the K^-1 product on top of each
    # other stay below 1e-11 (measured 2e-12), five orders inside what LML rtol 1e-8 / gradient 1e-6 need at this size
    assert float((L7 - Ld)[tril].abs().max()) <= 1e-11 * sl
    assert float((K7 - Kd)[tril].abs().max()) <= 1e-11 * sk
    # and both invert K: K * Kinv = I on a column sample
    Kfull = torch.tril(K7) + torch.tril(K7, -1).T
    cols = torch.arange(0, n, n // 16, device=engine.device)
    R = K @ Kfull[:, cols]
    R[cols, torch.arange(cols.numel(), device=engine.device)] -= 1.0
    assert float(R.abs().max()) <= 1e-9


@pytest.mark.parametrize("n", [4096, 8192])
def test_int8_trailing_updates_of_the_cholesky(engine, n):
    """Three-level blocked Cholesky (super-panels of 1024 columns whose rank-1024 trailing update runs on the int8 tensor
    pipe) against the all-DMMA factorisation and LAPACK; forced on here whatever the default is."""
    import torch
    torch.manual_seed(7)
    B = torch.randn(n, n // 2, dtype=torch.float64, device=engine.device)
    K = B @ B.T / n + 0.5 * torch.eye(n, dtype=torch.float64, device=engine.device)
    del B
    lib = engine.lib
    try:
        lib.mogp_set_i8(0, 7)
        Ad = K.clone()
        assert engine.potrf_(Ad) == 0
        lib.mogp_set_i8(2048, 7)
        lib.mogp_set_i8_potrf_min(4096)
        A7 = K.clone()
        assert engine.potrf_(A7) == 0
    finally:
        lib.mogp_set_i8(2048, 7)
        lib.mogp_set_i8_potrf_min(8192)
    ref = torch.linalg.cholesky(K)
    scale = float(ref.abs().max())
    assert float((torch.tril(Ad) - ref).abs().max()) <= 1e-11 * scale
    assert float((torch.tril(A7) - ref).abs().max()) <= 1e-11 * scale
    assert float((torch.tril(A7) - torch.tril(Ad)).abs().max()) <= 1e-11 * scale


def test_int8_path_survives_alternating_problem_sizes(engine):
    """One handle alternating between two sizes that both use the int8 path: the tile lists are rebuilt per size and the
    captured step graphs of the other size are invalidated (they would replay stale lists otherwise)."""
    ga, gb = load_golden("cfg4"), load_golden("cfg3")
    for g in (ga, gb, ga, gb, ga):
        for _ in range(3):                               # plain run, capture, replay
            res = engine.lml_grad(g["kind"], g["params"], g["sigma"], g["X"], g["y"], g["jitter"], True)
            assert abs(res["lml"] - float(g["lml"])) <= 1e-8 * abs(float(g["lml"]))


@pytest.mark.parametrize("leaf", [1024, 2048])
@pytest.mark.parametrize("n", [8192])
def test_recursive_cholesky_against_lapack(engine, n, leaf):
    """mogp_potrf through the recursive scheme (leaves by the blocked sweep, everything above on the int8 tensor pipe): factor
    against LAPACK, and the first bad pivot is reported with its global index whichever leaf it falls into."""
    lib = engine.lib
    assert lib.mogp_set_rchol(1, 4096, leaf) == 0
    try:
        gen = torch.Generator().manual_seed(n + leaf)
        B = torch.randn((n, n + 8), generator=gen, dtype=torch.float64)
        A = B @ B.T / n + 0.5 * torch.eye(n, dtype=torch.float64)
        Ad = A.cuda().clone()
        assert engine.potrf_(Ad) == 0
        L = torch.tril(Ad).cpu()
        Lref = torch.linalg.cholesky(A)
        assert float((L - Lref).abs().max() / Lref.abs().max()) < 1e-11
        for bad in (100, leaf + 77, n - 5):
            Ab = A.clone()
            Ab[bad, bad] = -1.0
            assert engine.potrf_(Ab.cuda()) == bad + 1
    finally:
        lib.mogp_set_rchol(1, 4096, 2048)


@pytest.mark.parametrize("rchol", [(0, 4096, 2048), (1, 4096, 2048), (1, 4096, 1024)])
@pytest.mark.parametrize("name", ["cfg4", "cfg3"])
def test_recursive_factor_and_inverse_in_the_step(engine, knobs, name, rchol):
    """The whole step at N = 4096 / 8192 with the recursive factor + inverse (on / off, leaf sizes): LML 1e-8, gradients 1e-6."""
    knobs()
    assert engine.lib.mogp_set_rchol(*rchol) == 0
    try:
        _check(engine, load_golden(name))
    finally:
        engine.lib.mogp_set_rchol(1, 4096, 2048)


@pytest.mark.parametrize("N,leaf", [(4352, 2176), (5120, 2560), (6144, 1536)])
def test_recursive_scheme_at_other_leaf_sizes(engine, knobs, N, leaf):
    """Padded sizes Np = leaf * 2^k with a leaf other than 2048 rows (any multiple of 128 in [1024, 2560]): the recursive factor +
    inverse against the blocked sweep on the same problem (two different schedules of the same mathematics)."""
    from mogptk_b200 import synth
    knobs()
    lib = engine.lib
    assert lib.mogp_set_rchol(1, 4096, 2048) == 0
    assert lib.mogp_rchol_leaf_for(N) == leaf
    X, y = synth.make_data(4, [N // 4] * 4, seed=5)
    p, sigma = synth.make_params("MOSM", 4, 3, 1, seed=5, random_delay_phase=True)
    res = {}
    try:
        for on in (1, 0):
            assert lib.mogp_set_rchol(on, 4096, 2048) == 0
            for _ in range(3):                       # plain run, capture, replay
                r = engine.lml_grad("MOSM", p, sigma, X, y, 1e-8, True)
            assert r["info"] == 0
            res[on] = r
    finally:
        lib.mogp_set_rchol(1, 4096, 2048)
    assert abs(res[1]["lml"] - res[0]["lml"]) <= 1e-10 * abs(res[0]["lml"])
    for k, g0 in res[0]["grad"].items():
        scale = max(float(g0.abs().max()), 1e-12)
        assert float((res[1]["grad"][k] - g0).abs().max()) <= 1e-8 * scale, k
