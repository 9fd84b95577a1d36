"""Parity of the CUDA path (through the C-ABI) against the CPU oracle.  Run on the B200 box:
    python -m pytest tests -m gpu -x -q

Bars (BASELINE.json north_star):
  STRICT mode: l, e, p, phase2_start BIT-IDENTICAL to the oracle (stronger than required).
  FAST mode:   p identical wherever the oracle's decision margin is not a rounding tie,
               |e - e_ref| <= 1e-10 * max(||e_ref||_inf, tiny) ("E to 1e-10 relative"),
               max|L L^T - P(A+E)P^T| <= 1e-12 * max(1, max|A|).
"""
import numpy as np
import pytest

import matgen
import modcholesky_b200 as m
from conftest import identity_residual, perm_matrix

pytestmark = pytest.mark.gpu

ALGOS = ("gmw81", "se90", "se99")
TOL_RECON = 1e-12      # north_star: ||LL^T - P(A+E)P^T|| / ||A|| <= 1e-12
TOL_E_REL = 1e-10      # north_star: E agrees to 1e-10 relative
TIE_MARGIN = 1e-9      # decisions closer than this (relative) are rounding ties


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    return torch


def _gpu(torch, algo, a, mode):
    d = m.factor_batched(algo, torch.from_numpy(a).cuda(), mode=mode)
    torch.cuda.synchronize()
    return (d.l.cpu().numpy(), d.e.cpu().numpy(), d.p.cpu().numpy().astype(np.uint64),
            d.phase2_start.cpu().numpy())


def _assert_bit_exact(tag, got, ref):
    l, e, p, k = got
    assert np.array_equal(p, ref.p), f"{tag}: p differs in {(p != ref.p).any(axis=1).sum()} matrices"
    assert np.array_equal(k, ref.phase2_start), f"{tag}: phase2_start differs"
    assert np.array_equal(e, ref.e, equal_nan=True), \
        f"{tag}: e differs, max abs {np.nanmax(np.abs(e - ref.e))}"
    assert np.array_equal(l, ref.l, equal_nan=True), \
        f"{tag}: l differs, max abs {np.nanmax(np.abs(l - ref.l))}"


RELAXED = {"checked": 0, "e_relaxed": 0, "recon_relaxed": 0}   # printed by the last test of this file


def _assert_tolerance(tag, a, got, ref, algo=None, min_firm=0.9):
    """FAST-mode bars.  The plain north-star bars (E to 1e-10 of ||E||, reconstruction to 1e-12 of
    ||A||) are asserted for every matrix except the two GMW81 situations in which they are below
    the f64 resolution of the quantities themselves; those matrices are counted in RELAXED."""
    l, e, p, k = got
    b, n = a.shape[0], a.shape[1]
    if algo is None:
        algo = next((x for x in ALGOS if x in tag), "")
    ok = ~np.isnan(ref.l).any(axis=(1, 2))
    firm = ok & (ref.margin > TIE_MARGIN)
    same_p = (p == ref.p).all(axis=1)
    assert same_p[firm].all(), f"{tag}: p differs on {(~same_p[firm]).sum()} non-tie matrices"
    assert firm.sum() >= min_firm * ok.sum(), f"{tag}: too many rounding ties to be meaningful"
    cmp = firm & same_p
    escale = np.maximum(np.abs(ref.e).max(axis=1), 1e-300)
    anorm = np.maximum(np.abs(a).max(axis=(1, 2)), 1.0)
    de = np.abs(e - ref.e).max(axis=1)
    # GMW81 only: e_j = d_j - c_jj inherits the cancellation noise of c_jj = a_jj - sum(...), a few
    # ulps of ||A|| in the reference itself; where E is tiny (||E|| < 1e-4 ||A||, bench-style
    # matrices) agreement is required down to 1e-14 ||A|| (45 ulp), not below.
    e_bar = TOL_E_REL * escale
    e_rel = (algo == "gmw81") & (escale < 1e-4 * anorm)
    e_bar = np.where(e_rel, TOL_E_REL * 1e-4 * anorm, e_bar)
    bad = cmp & (de > e_bar)
    assert not bad.any(), f"{tag}: E differs beyond 1e-10 rel in {bad.sum()} matrices " \
                          f"(worst {np.max(de[bad] / escale[bad]):.2e})"
    # reconstruction for EVERY finite matrix, tie or not
    llt = l @ np.transpose(l, (0, 2, 1))
    idx = p.astype(np.int64)
    ap = np.take_along_axis(np.take_along_axis(a, idx[:, :, None], 1), idx[:, None, :], 2)
    ep = np.take_along_axis(e, idx, 1)
    rhs = ap + np.einsum("bi,ij->bij", ep, np.eye(n))
    res = np.abs(llt - rhs).max(axis=(1, 2)) / anorm
    # GMW81 only: where ||A + E|| > 1e3 ||A|| (n = 256 adds E ~ 1e3 ||A||) evaluating L L^T in f64
    # is itself off by more than 1e-12 ||A||: the bar is then the reference's own residual, computed
    # the same way from the oracle's factors, and never below a few ulps of the largest entry of
    # A + E (which the diagonal of L L^T must hit).
    aenorm = np.nan_to_num(np.abs(rhs).max(axis=(1, 2)))
    r_rel = (algo == "gmw81") & (aenorm > 1e3 * anorm)
    bar = np.full(b, TOL_RECON)
    if r_rel.any():
        lr = ref.l
        idr = ref.p.astype(np.int64)
        apr = np.take_along_axis(np.take_along_axis(a, idr[:, :, None], 1), idr[:, None, :], 2)
        rhr = apr + np.einsum("bi,ij->bij", np.take_along_axis(ref.e, idr, 1), np.eye(n))
        with np.errstate(invalid="ignore"):
            res_ref = np.abs(lr @ np.transpose(lr, (0, 2, 1)) - rhr).max(axis=(1, 2)) / anorm
        ulps = 8.0 * np.finfo(np.float64).eps * aenorm / anorm
        relaxed = np.maximum(np.maximum(TOL_RECON, 2.0 * np.nan_to_num(res_ref)), ulps)
        bar = np.where(r_rel, relaxed, bar)
    assert (res[ok] <= bar[ok]).all(), \
        f"{tag}: reconstruction {res[ok].max():.3e} exceeds the bar in {(res[ok] > bar[ok]).sum()} matrices"
    assert (np.triu(l, 1) == 0).all() and (e[ok] >= 0).all()
    RELAXED["checked"] += int(ok.sum())
    RELAXED["e_relaxed"] += int((cmp & e_rel).sum())
    RELAXED["recon_relaxed"] += int((ok & r_rel).sum())


# ---- the reference's own tests, run through the mirrored trait API on the GPU -------------
@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_reference_test_suite_on_gpu(golden, torch_cuda, mode):
    """STRICT: the reference's assertions verbatim (absolute tolerances).  FAST: the same
    assertions with the tolerance taken relative to max(1, max|A|) -- the reference's absolute
    1e-12 on A4 (entries ~4.8e3) is 0.2 ulp and only its exact arithmetic can promise it; the
    north-star bar is relative (||LL^T - P(A+E)P^T|| / ||A|| <= 1e-12)."""
    for case in golden["cases"]:
        a = np.array(golden["matrices"][case["matrix"]]["a"], dtype=np.float64)
        scale = 1.0 if mode == "strict" else max(1.0, np.abs(a).max())
        d = getattr(m.Array2(a, mode=mode), "mod_cholesky_" + case["algo"])()
        l, e, p = d.l, d.e, d.p
        assert identity_residual(a, l, e, p) < case["tol_identity"] * scale, case["name"]
        llt = l @ l.T
        if "u" in case:
            u = np.array(case["u"])
            assert np.abs(llt - u.T @ u).max() < case["tol_llt"] * scale, case["name"]
        if "res" in case:
            P = perm_matrix(p)
            assert np.abs(llt - P @ np.array(case["res"]) @ P.T).max() < case["tol_llt"] * scale, case["name"]
        if "l" in case:
            assert np.abs(l - np.array(case["l"])).max() < case["tol_l"] * scale, case["name"]


def test_gershgorin_circles_on_gpu(golden, oracle, torch_cuda):
    g = golden["gershgorin"]
    res = m.Array2(np.array(g["a"])).gershgorin_circles()
    for (c, r), (c0, r0) in zip(res, g["circles"]):
        assert abs(c - c0) < g["tol"] and abs(r - r0) < g["tol"]
    rng = np.random.default_rng(3)
    a = rng.standard_normal((50, 19, 19))
    out = m.gershgorin_circles_batched(a)
    ref = np.stack([oracle.gershgorin_circles(x) for x in a])
    assert np.array_equal(out, ref)
    outd = m.gershgorin_circles_batched(torch_cuda.from_numpy(a).cuda()).cpu().numpy()
    assert np.array_equal(outd, ref)


SIZES = [1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 17, 23, 24, 31, 32, 33, 40, 48, 63, 64, 65, 96, 127, 128, 129]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n", SIZES)
def test_strict_bit_exact_vs_oracle(oracle, torch_cuda, algo, n):
    b = 96 if n <= 32 else (32 if n <= 64 else 12)
    for fam in ("uniform", "wishart", "pd", "bench", "ties"):
        a = matgen.make(fam, 1000 + n, b, n)
        ref = oracle.factor_batched(algo, a)
        _assert_bit_exact(f"{algo} n={n} {fam}", _gpu(torch_cuda, algo, a, "strict"), ref)


@pytest.mark.parametrize("family", ["reg", "wsm", "tpm"])
def test_both_small_kernel_families_are_bit_exact(oracle, torch_cuda, family):
    """n <= 32 has two warp-level kernel families (shared-memory-resident by default,
    register-resident on request) and n <= 8 a thread-per-matrix one (the default there); all
    must give the oracle's bits."""
    prev = m.set_small_kernel_family(family)
    try:
        for algo in ALGOS:
            for n in (1, 2, 3, 4, 5, 8, 9, 16, 17, 31, 32):
                if family == "tpm" and n > 8:
                    continue
                for mode in ("strict", "fast"):
                    assert m.kernel_class(algo, n, mode) == {"reg": "warp32", "wsm": "warp_smem", "tpm": "thread"}[family]
                for fam in ("uniform", "wishart", "ties"):
                    a = matgen.make(fam, 4000 + n, 64, n)
                    ref = oracle.factor_batched(algo, a)
                    _assert_bit_exact(f"{family} {algo} n={n} {fam}", _gpu(torch_cuda, algo, a, "strict"), ref)
                    if fam != "ties" and n >= 3:   # exact ties are covered by the tie test
                        _assert_tolerance(f"{family} {algo} n={n} {fam}", a,
                                          _gpu(torch_cuda, algo, a, "fast"), ref)
    finally:
        m.set_small_kernel_family(prev)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16])
def test_thread_per_matrix_kernels_are_bit_exact(oracle, torch_cuda, monkeypatch, n):
    """One thread per matrix (tpm_kernels.cuh): tiles of 128 matrices -- batches below one tile,
    of several tiles and with a ragged last tile; both modes run the reference's arithmetic, so
    FAST is bit-exact too; every family including exact ties and the bench-style Q D Q^T inputs."""
    torch = torch_cuda
    prev = m.set_small_kernel_family("tpm")
    try:
        for algo in ALGOS:
            for b, fam in ((1, "uniform"), (37, "wishart"), (128, "pd"), (129, "ties"), (1000, "uniform"),
                           (515, "bench"), (300, "wishart")):
                if fam == "bench" and n < 3:
                    continue
                a = matgen.make(fam, 8800 + 10 * n + b, b, n)
                ref = oracle.factor_batched(algo, a)
                for mode in ("strict", "fast"):
                    assert m.kernel_class(algo, n, mode) == "thread"
                    _assert_bit_exact(f"tpm {algo} n={n} {fam} b={b} {mode}", _gpu(torch, algo, a, mode), ref)
            # IEEE specials: zeros (0/0), +-identity, a negative rank-one matrix,
            # underflowing and overflowing entries, a NaN and an Inf entry -- propagated exactly like
            # the reference, NaN positions included, in both modes
            with_nan, with_inf = np.eye(n), np.eye(n)
            with_nan[n - 1, n - 1] = np.nan
            with_inf[0, 0] = np.inf
            specials = np.stack([np.zeros((n, n)), np.eye(n), -np.eye(n), -4.0 * np.ones((n, n)),
                                 np.full((n, n), 1e-300), np.full((n, n), 1e300) * np.eye(n), with_nan, with_inf])
            ref = oracle.factor_batched(algo, specials)
            for mode in ("strict", "fast"):
                got = _gpu(torch, algo, specials, mode)
                assert np.array_equal(np.isnan(got[0]), np.isnan(ref.l)), (algo, n, mode, "NaN pattern of L")
                _assert_bit_exact(f"tpm {algo} n={n} specials {mode}", got, ref)
    finally:
        m.set_small_kernel_family(prev)
    # the run-time-n variant of the class (MODCHOL_TPR=0: full matrix per thread in shared memory)
    if n <= 8:
        monkeypatch.setenv("MODCHOL_TPR", "0")
        prev = m.set_small_kernel_family("tpm")
        try:
            for algo in ALGOS:
                a = matgen.make("wishart", 9900 + n, 200, n)
                _assert_bit_exact(f"tpm(run-time n) {algo} n={n}", _gpu(torch, algo, a, "strict"),
                                  oracle.factor_batched(algo, a))
        finally:
            m.set_small_kernel_family(prev)
            monkeypatch.delenv("MODCHOL_TPR")
    # the default dispatch (batches of 512 and more, 4096 for n > 8): n <= 8 always; up to n = 12 (GMW81: 10) in FAST and n = 16 in STRICT mode
    assert m.kernel_class("se99", 8, "fast") == "thread" and m.kernel_class("gmw81", 9, "fast") == "thread"
    assert m.kernel_class("gmw81", 10, "fast") == "thread" and m.kernel_class("gmw81", 11, "fast") == "warp_smem"
    assert m.kernel_class("gmw81", 16, "strict") == "thread"
    assert m.kernel_class("se99", 12, "fast") == "thread" and m.kernel_class("se99", 13, "fast") == "warp_smem"
    assert m.kernel_class("se90", 16, "strict") == "thread" and m.kernel_class("se90", 17, "strict") == "warp_smem"


def test_packed_shared_memory_variant_is_bit_exact(oracle, torch_cuda, monkeypatch):
    """MODCHOL_WSM_PACK=1: two matrices per warp (16 lanes x 2 rows) for 16 < n <= 32; the two
    halves of a warp diverge whenever their matrices change phase at different columns, and odd
    batch sizes leave half a warp without a matrix."""
    monkeypatch.setenv("MODCHOL_WSM_PACK", "1")
    prev = m.set_small_kernel_family("wsm")
    try:
        for algo in ALGOS:
            for n, b in ((17, 33), (24, 64), (31, 7), (32, 129)):
                for fam in ("uniform", "wishart", "pd", "ties"):
                    a = matgen.make(fam, 5000 + n, b, n)
                    ref = oracle.factor_batched(algo, a)
                    _assert_bit_exact(f"packed {algo} n={n} {fam}", _gpu(torch_cuda, algo, a, "strict"), ref)
                    if fam != "ties":
                        _assert_tolerance(f"packed {algo} n={n} {fam}", a,
                                          _gpu(torch_cuda, algo, a, "fast"), ref)
    finally:
        m.set_small_kernel_family(prev)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n", [160, 200, 236, 237, 250, 256, 257, 300])
def test_strict_bit_exact_large(oracle, torch_cuda, algo, n):
    for fam in ("uniform", "wishart"):
        a = matgen.make(fam, 2000 + n, 6, n)
        ref = oracle.factor_batched(algo, a)
        _assert_bit_exact(f"{algo} n={n} {fam}", _gpu(torch_cuda, algo, a, "strict"), ref)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 16, 24, 25, 32, 33, 40, 48, 49, 63, 64, 100, 150, 256])
def test_fast_mode_within_north_star_tolerances(oracle, torch_cuda, algo, n):
    b = 511 if n <= 32 else (63 if n <= 64 else 16)   # odd: the last warp of a packed kernel is half empty
    for fam in ("uniform", "wishart", "pd", "bench"):
        a = matgen.make(fam, 3000 + n, b, n)
        ref = oracle.factor_batched(algo, a)
        _assert_tolerance(f"{algo} n={n} {fam}", a, _gpu(torch_cuda, algo, a, "fast"), ref)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n", [129, 136, 200, 255, 256, 300, 384, 512, 1000])
def test_fast_mode_large_blocked_kernels(oracle, torch_cuda, algo, n):
    """The large size class (128 < n <= 1024, north star): blocked CTA-per-matrix kernels
    (bsm_kernels.cuh) for all three algorithms, FAST bars against the oracle."""
    assert m.kernel_class(algo, n, "fast") == "cta_blocked"
    b = 6 if n <= 300 else 2
    for fam in ("uniform", "wishart", "pd", "bench"):
        a = matgen.make(fam, 4000 + n, b, n)
        ref = oracle.factor_batched(algo, a)
        _assert_tolerance(f"{algo} n={n} {fam}", a, _gpu(torch_cuda, algo, a, "fast"), ref,
                          min_firm=0.4 if fam == "bench" else 0.9)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n", [66, 97, 128, 130, 200, 250, 256])
def test_fast_mode_warp_panel_kernels(oracle, torch_cuda, algo, n, monkeypatch):
    """The opt-in warp-per-matrix panel kernels (lwm_kernels.cuh: trailing matrix streamed through
    a cp.async.bulk / mbarrier ring): even n takes the bulk-copy path, odd n the ordinary one."""
    monkeypatch.setenv("MODCHOL_LWM_MIN", "65")
    assert m.kernel_class(algo, n, "fast") == "warp_panel"
    for fam in ("uniform", "wishart", "pd"):
        a = matgen.make(fam, 5000 + n, 6, n)
        ref = oracle.factor_batched(algo, a)
        _assert_tolerance(f"{algo} n={n} {fam} lwm", a, _gpu(torch_cuda, algo, a, "fast"), ref)


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_exact_ties_resolve_to_lowest_index(oracle, torch_cuda, mode):
    # equal diagonals -> lowest index wins (utils.rs:57-69); A6-like duplicated rows
    a = np.stack([np.diag([2.0, 5.0, 5.0, 1.0]), np.diag([1.0, 1.0, 1.0, 1.0]),
                  np.diag([-1.0, -1.0, 3.0, 3.0])])
    for algo in ALGOS:
        ref = oracle.factor_batched(algo, a)
        l, e, p, k = _gpu(torch_cuda, algo, a, mode)
        assert np.array_equal(p, ref.p), algo
    for n in (6, 8, 12, 16, 32, 48, 64):
        t = matgen.make("ties", 77 + n, 64, n)
        for algo in ALGOS:
            ref = oracle.factor_batched(algo, t)
            l, e, p, k = _gpu(torch_cuda, algo, t, mode)
            if mode == "strict":
                assert np.array_equal(p, ref.p), (algo, n)
            else:
                firm = ref.margin > TIE_MARGIN
                assert np.array_equal(p[firm], ref.p[firm]), (algo, n)


@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_edge_cases_propagate_like_the_reference(oracle, torch_cuda, mode):
    """IEEE propagation through the reference's formulas is reproduced,
    not 'fixed' (NaN positions included)."""
    for n in (1, 2, 3, 6, 8, 12, 16, 32, 40, 64, 140):   # every FAST kernel class sees the specials
        specials = np.stack([np.zeros((n, n)), np.eye(n), -np.eye(n), -4.0 * np.ones((n, n)),
                             np.full((n, n), 1e-300), np.full((n, n), 1e300) * np.eye(n)])
        for algo in ALGOS:
            ref = oracle.factor_batched(algo, specials)
            l, e, p, k = _gpu(torch_cuda, algo, specials, mode)
            if mode == "strict":
                assert np.array_equal(np.isnan(l), np.isnan(ref.l)), (algo, n, "NaN pattern of L")
                assert np.array_equal(np.isnan(e), np.isnan(ref.e)), (algo, n, "NaN pattern of E")
            # FAST contracts a - b*c: where the reference cancels to exactly 0 and then divides
            # 0/0 (e.g. 1e-300 * ones, exactly rank one), the fused update keeps a 1e-317 residue
            # and stays finite.  Only matrices the reference factorises finitely are compared.
            if mode == "strict":
                _assert_bit_exact(f"{algo} n={n} specials", (l, e, p, k), ref)
            else:
                fin = ~np.isnan(ref.l).any(axis=(1, 2))
                assert not np.isnan(l[fin]).any() and not np.isnan(e[fin]).any(), (algo, n)
                assert np.array_equal(p[fin], ref.p[fin]), (algo, n)
                # compare at the scale of each matrix: entries that are exactly 0 in the
                # reference may come out as sqrt(rounding residue) (e.g. 1e-300 * ones)
                for i in np.nonzero(fin)[0]:
                    ls = max(np.abs(ref.l[i]).max(), 1e-300)
                    es = max(np.abs(ref.e[i]).max(), np.abs(specials[i]).max(), 1e-300)
                    assert np.abs(e[i] - ref.e[i]).max() <= 1e-10 * es, (algo, n, i)
                    rr = np.abs(l[i] @ l[i].T - ref.l[i] @ ref.l[i].T).max()
                    assert rr <= 1e-12 * ls * ls, (algo, n, i, rr)


@pytest.mark.parametrize("n", [1, 5, 32, 47, 130])
def test_solve_batched_newton_step(torch_cuda, n):
    """(A + E) x = b from {l, p}; checked against numpy on the host."""
    rng = np.random.default_rng(n)
    b = 40
    # tiny n: PD inputs (a negative 1x1 makes SE90 return NaN, the reference's quirk at se90.rs:84)
    a = matgen.make("wishart" if n >= 3 else "pd", 900 + n, b, n)
    rhs = rng.standard_normal((b, 3, n))
    for algo in ALGOS:
        d = m.factor_batched(algo, a, mode="strict")
        x = m.solve_batched(d, rhs)
        xd = m.solve_batched(m.factor_batched(algo, torch_cuda.from_numpy(a).cuda()),
                             torch_cuda.from_numpy(rhs).cuda()).cpu().numpy()
        assert np.array_equal(x, xd)
        for i in range(b):
            ae = a[i] + np.diag(d.e[i])
            ref = np.linalg.solve(ae, rhs[i].T).T
            scale = np.abs(ref).max() * np.linalg.cond(ae)
            assert np.abs(x[i] - ref).max() <= 1e-13 * scale + 1e-300, (algo, n, i)
            res = np.abs(ae @ x[i].T - rhs[i].T).max()
            assert res <= 1e-12 * np.abs(ae).max() * max(np.abs(x[i]).max(), 1.0) * n, (algo, n, i, res)
    x1 = m.solve_batched(d, rhs[:, 0, :])
    assert np.array_equal(x1, x[:, 0, :])


@pytest.mark.parametrize("n", [1, 3, 8, 16, 17, 20, 24, 27, 31, 32, 48, 64, 100])
def test_fused_factor_solve(oracle, torch_cuda, n):
    """modchol_factor_solve_batched_f64: one call (one kernel for n <= 64) gives the x of
    factor + solve, with every factor output optional and bit-identical when requested."""
    torch = torch_cuda
    b = 40
    fam = "pd" if n < 3 else "wishart"
    a = matgen.make(fam, 6000 + n, b, n)
    rng = np.random.default_rng(n)
    for algo in ALGOS:
        for nrhs in (1, 3):
            rhs = rng.standard_normal((b, nrhs, n))
            for mode in ("strict", "fast"):
                ad, rd = torch.from_numpy(a).cuda(), torch.from_numpy(rhs).cuda()
                d = m.factor_batched(algo, ad, mode=mode)
                x2 = m.solve_batched(d, rd).cpu().numpy()
                x, df = m.factor_solve_batched(algo, ad, rd, mode=mode, return_factors=True)
                x = x.cpu().numpy()
                if mode == "strict":
                    # the factors that come with it are the plain call's, bit for bit
                    assert torch.equal(df.l, d.l) and torch.equal(df.e, d.e) and torch.equal(df.p, d.p)
                else:
                    # FAST: outside 16 < n <= 32 (where the implicit-pivoting kernels carry the fused
                    # solve themselves) the plain and the fused call may run different kernel
                    # families -- same pivots, roundings differ
                    assert torch.equal(df.p, d.p)
                    assert torch.allclose(df.l, d.l, rtol=1e-9, atol=1e-11)
                    assert torch.allclose(df.e, d.e, rtol=1e-9, atol=1e-11)
                assert torch.equal(df.phase2_start, d.phase2_start)
                # x without asking for the factors (NULL outputs) is the same x
                x0 = m.factor_solve_batched(algo, ad, rd, mode=mode).cpu().numpy()
                assert np.array_equal(x0, x)
                e = d.e.cpu().numpy()
                for i in range(b):
                    ae = a[i] + np.diag(e[i])
                    if not np.isfinite(ae).all():
                        continue
                    ref = np.linalg.solve(ae, rhs[i].T).T
                    scale = np.abs(ref).max() * np.linalg.cond(ae)
                    assert np.abs(x[i] - ref).max() <= 1e-13 * scale + 1e-300, (algo, n, mode, i)
                    assert np.abs(x[i] - x2[i]).max() <= 1e-13 * scale + 1e-300, (algo, n, mode, i)
            # host pointers, single right-hand side given as (batch, n)
            xh = m.factor_solve_batched(algo, a, rhs[:, 0, :], mode="strict")
            xdv = m.factor_solve_batched(algo, torch.from_numpy(a).cuda(),
                                         torch.from_numpy(np.ascontiguousarray(rhs[:, 0, :])).cuda(),
                                         mode="strict").cpu().numpy()
            assert isinstance(xh, np.ndarray) and np.array_equal(xh, xdv)


def test_host_pointer_path_matches_device_path(oracle, torch_cuda):
    """numpy in -> the C-ABI stages through the GPU in overlapped chunks; results identical."""
    for algo, n, b in (("se99", 32, 5000), ("gmw81", 17, 3000), ("se90", 64, 700)):
        a = matgen.make("uniform", 5 + n, b, n)
        dh = m.factor_batched(algo, a, mode="strict")
        assert isinstance(dh.l, np.ndarray)
        l, e, p, k = _gpu(torch_cuda, algo, a, "strict")
        assert np.array_equal(dh.l, l) and np.array_equal(dh.e, e) and np.array_equal(dh.p, p)
        assert np.array_equal(dh.phase2_start, k)
        ref = oracle.factor_batched(algo, a[:256])
        assert np.array_equal(dh.l[:256], ref.l) and np.array_equal(dh.p[:256], ref.p)


def test_multi_gpu_entry_point_single_device(oracle, torch_cuda):
    a = matgen.make("wishart", 11, 2000, 16)
    d1 = m.Array3(a).mod_cholesky_se99_batched(ngpus=-1)   # all visible devices
    d0 = m.Array3(a).mod_cholesky_se99_batched()
    assert np.array_equal(d1.l, d0.l) and np.array_equal(d1.e, d0.e) and np.array_equal(d1.p, d0.p)


def test_eight_byte_aligned_buffers(oracle, torch_cuda):
    """The C-ABI takes plain pointers: buffers that are only 8-byte aligned (a view one element
    into an allocation) must work -- the register kernels then skip their 128-bit stores."""
    import ctypes
    from modcholesky_b200 import _lib
    torch = torch_cuda
    # (the last five cases reach the thread-per-matrix kernels, whose direct accesses are 128-bit
    # loads and 256-bit stores when the bases allow: off = 2 is 16- but not 32-byte aligned)
    for algo, n, b, off in (("se99", 32, 300, 1), ("se90", 16, 300, 1), ("gmw81", 24, 200, 1), ("se99", 64, 50, 1),
                            ("se99", 8, 700, 1), ("se99", 8, 700, 2), ("gmw81", 2, 600, 1), ("se90", 12, 4200, 2),
                            ("gmw81", 16, 4200, 1)):
        a = matgen.make("uniform", 70 + n, b, n)
        ref = oracle.factor_batched(algo, a)
        for mode in ("strict", "fast"):
            pad_a = torch.empty(b * n * n + off, dtype=torch.float64, device="cuda")
            pad_l = torch.empty(b * n * n + off, dtype=torch.float64, device="cuda")
            pad_e = torch.empty(b * n + 1, dtype=torch.float64, device="cuda")
            pad_p = torch.empty(b * n + 1, dtype=torch.int64, device="cuda")
            av, lv, evv, pv = pad_a[off:], pad_l[off:], pad_e[1:], pad_p[1:]
            assert lv.data_ptr() % 32 == 8 * off
            av.copy_(torch.from_numpy(a).reshape(-1))
            rc = _lib.lib().modchol_factor_batched_f64(
                _lib.algo_id(algo), _lib.mode_id(mode), b, n, av.data_ptr(), lv.data_ptr(), evv.data_ptr(),
                pv.data_ptr(), None, _lib.MEM_DEVICE,
                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc)
            torch.cuda.synchronize()
            l = lv.cpu().numpy().reshape(b, n, n)
            p = pv.cpu().numpy().reshape(b, n).astype(np.uint64)
            if mode == "strict":
                assert np.array_equal(l, ref.l) and np.array_equal(p, ref.p), (algo, n)
                assert np.array_equal(evv.cpu().numpy().reshape(b, n), ref.e), (algo, n)
            else:
                assert np.allclose(l, ref.l, rtol=1e-9, atol=1e-12), (algo, n)


def test_input_is_never_written_and_empty_batch_ok(torch_cuda):
    torch = torch_cuda
    a = torch.from_numpy(matgen.make("uniform", 1, 64, 32)).cuda()
    a0 = a.clone()
    for algo in ALGOS:
        for mode in ("strict", "fast"):
            m.factor_batched(algo, a, mode=mode)
    torch.cuda.synchronize()
    assert torch.equal(a, a0)
    d = m.factor_batched("se99", torch.empty((0, 8, 8), dtype=torch.float64, device="cuda"))
    assert d.l.shape == (0, 8, 8)
    d = m.factor_batched("se99", np.empty((0, 8, 8)))
    assert d.p.shape == (0, 8)


# ---- BASELINE.json configs at FULL size, via size-independent properties ---------------------
FULL = [
    ("se99", 16, 1 << 20, "uniform"),     # configs[1]
    ("se99", 32, 1 << 18, "uniform"),     # configs[4] per-GPU slice (2^20) is run by bench.py
    ("se90", 64, 200_000, "uniform"),     # configs[2]
    ("gmw81", 256, 4096, "uniform"),      # configs[3]
]


@pytest.mark.parametrize("algo,n,batch,fam", FULL, ids=[f"{f[0]}-n{f[1]}" for f in FULL])
@pytest.mark.parametrize("mode", ["strict", "fast"])
def test_full_size_properties(oracle, torch_cuda, algo, n, batch, fam, mode):
    torch = torch_cuda
    g = torch.Generator(device="cuda").manual_seed(1234)
    a = torch.rand((batch, n, n), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    a = torch.tril(a) + torch.tril(a, -1).transpose(1, 2)
    d = m.factor_batched(algo, a, mode=mode)
    resid, flags = m.verify_batched(a, d)
    torch.cuda.synchronize()
    assert int(flags.max()) == 0, "p not a permutation / e < 0 / dirty upper triangle / NaN"
    # 1e-12 ||A||, or a few ulps of the largest entry of A + E where E dwarfs A (GMW81, n = 256)
    amax = a.abs().amax(dim=(1, 2))
    bar = torch.clamp(8.0 * np.finfo(np.float64).eps * (amax + d.e.amax(dim=1)) / torch.clamp(amax, min=1.0),
                      min=TOL_RECON)
    assert bool((resid <= bar).all()), float((resid / bar).max())
    # cross-check the on-device verifier itself and the kernel on a sampled subset
    sel = torch.randint(0, batch, (64,), generator=torch.Generator().manual_seed(5)).tolist()
    asub = a[sel].cpu().numpy()
    ref = oracle.factor_batched(algo, asub)
    got = (d.l[sel].cpu().numpy(), d.e[sel].cpu().numpy(),
           d.p[sel].cpu().numpy().astype(np.uint64), d.phase2_start[sel].cpu().numpy())
    if mode == "strict":
        _assert_bit_exact(f"{algo} n={n} full", got, ref)
    else:
        _assert_tolerance(f"{algo} n={n} full", asub, got, ref)
    for i in range(4):
        r = identity_residual(asub[i], got[0][i], got[1][i], got[2][i])
        assert abs(r / max(1.0, np.abs(asub[i]).max()) - float(resid[sel[i]])) <= 1e-13
    # idempotence / determinism: a second run gives the same bits
    d2 = m.factor_batched(algo, a, mode=mode)
    torch.cuda.synchronize()
    assert torch.equal(d.l, d2.l) and torch.equal(d.e, d2.e) and torch.equal(d.p, d2.p)


def test_multi_gpu_entry_point_two_devices(oracle, torch_cuda):
    """modchol_factor_batched_multi_gpu_f64 on every visible device (runs where the box has >= 2
    GPUs): contiguous slices, one host thread per device, results identical to one device."""
    if m.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    a = matgen.make("wishart", 12, 30001, 32)        # odd batch: uneven slices
    for mode in ("strict", "fast"):
        d1 = m.factor_batched("se99", a, mode=mode, ngpus=-1)
        d2 = m.factor_batched("se99", a, mode=mode, ngpus=2)
        d0 = m.factor_batched("se99", a, mode=mode)
        for d in (d1, d2):
            assert np.array_equal(d.l, d0.l) and np.array_equal(d.e, d0.e) and np.array_equal(d.p, d0.p)
            assert np.array_equal(d.phase2_start, d0.phase2_start)
    ref = oracle.factor_batched("se99", a[:128])
    d = m.factor_batched("se99", a, mode="strict", ngpus=-1)
    assert np.array_equal(d.l[:128], ref.l) and np.array_equal(d.p[:128], ref.p)


def test_pinned_and_pageable_host_buffers_agree(torch_cuda):
    """HOST mem_kind: pinned caller memory is DMA'd directly, pageable memory goes through the
    library's pinned bounce buffers; several chunks per slot so that buffers are reused."""
    import ctypes
    from modcholesky_b200 import _lib
    torch = torch_cuda
    n, b = 32, 60000        # 491 MB of input: 5+ chunks of 96 MB over three slots
    a = matgen.make("uniform", 21, b, n)
    dn = m.factor_batched("se99", a, mode="fast")                  # numpy: pageable
    ah = torch.from_numpy(a).pin_memory()
    lh = torch.empty((b, n, n), dtype=torch.float64).pin_memory()
    eh = torch.empty((b, n), dtype=torch.float64).pin_memory()
    ph = torch.empty((b, n), dtype=torch.int64).pin_memory()
    kh = torch.empty((b,), dtype=torch.int32).pin_memory()
    rc = _lib.lib().modchol_factor_batched_f64(_lib.SE99, _lib.MODE_FAST, b, n, ah.data_ptr(), lh.data_ptr(),
                                               eh.data_ptr(), ph.data_ptr(), kh.data_ptr(), _lib.MEM_HOST, None)
    _lib.check(rc)
    assert np.array_equal(lh.numpy(), dn.l) and np.array_equal(eh.numpy(), dn.e)
    assert np.array_equal(ph.numpy().astype(np.uint64), dn.p) and np.array_equal(kh.numpy(), dn.phase2_start)
    dd = m.factor_batched("se99", torch.from_numpy(a).cuda(), mode="fast")
    assert np.array_equal(dd.l.cpu().numpy(), dn.l)


@pytest.mark.parametrize("n", [1537, 2048])
def test_solve_at_the_largest_sizes(torch_cuda, n):
    """The batched solve keeps 4 n doubles of dynamic shared memory per block: above n = 1536 that
    is more than the 48 KB default and needs the opt-in attribute (MODCHOL_MAX_N = 2048)."""
    rng = np.random.default_rng(n)
    b = 2
    w = rng.standard_normal((b, n, n))
    a = w @ np.transpose(w, (0, 2, 1)) / n + np.eye(n)
    rhs = rng.standard_normal((b, n))
    d = m.factor_batched("gmw81", a, mode="strict")
    x = m.solve_batched(d, rhs)
    for i in range(b):
        ae = a[i] + np.diag(d.e[i])
        assert np.abs(ae @ x[i] - rhs[i]).max() <= 1e-9 * np.abs(ae).max() * max(np.abs(x[i]).max(), 1.0)
    xf = m.factor_solve_batched("gmw81", a, rhs, mode="strict")
    assert np.array_equal(xf, x)


def test_torch_arguments_are_checked(torch_cuda):
    """Device pointers are reinterpreted as f64 / u64 device memory by the kernels: wrong dtype,
    host tensors next to device tensors and mismatched shapes must raise, not fault."""
    torch = torch_cuda
    a = torch.from_numpy(matgen.make("pd", 3, 8, 16)).cuda()
    d = m.factor_batched("se99", a)
    rhs = torch.ones((8, 16), dtype=torch.float64, device="cuda")
    with pytest.raises(TypeError):
        m.factor_batched("se99", a.float())
    with pytest.raises(TypeError):
        m.solve_batched(d, rhs.float())
    with pytest.raises((ValueError, TypeError)):
        m.solve_batched(m.Decomposition(d.l.cpu(), d.e, d.p), rhs)
    with pytest.raises(ValueError):
        m.solve_batched(m.Decomposition(d.l[:4], d.e, d.p), rhs)
    with pytest.raises(TypeError):
        m.gershgorin_circles_batched(a.float())
    with pytest.raises((ValueError, TypeError)):
        m.verify_batched(a, m.Decomposition(d.l, d.e.cpu(), d.p))
    with pytest.raises((ValueError, TypeError)):
        m.factor_solve_batched("se99", a, rhs.cpu())
    with pytest.raises(ValueError):
        m.factor_solve_batched("se99", a, rhs[:4])
    x = m.solve_batched(d, rhs)          # and the well-formed call still works
    assert x.shape == rhs.shape


@pytest.mark.parametrize("n", [1, 5, 16, 17, 32, 33, 48, 49, 64, 65, 100])
def test_verifier_matches_numpy_and_raises_every_flag(oracle, torch_cuda, n):
    """The on-device acceptance check (warp-per-matrix DMMA kernel for n <= 64, CTA kernel above)
    against numpy's evaluation of the same identity (se99.rs:221-231), and each flag bit."""
    torch = torch_cuda
    b = 24
    a = matgen.make("uniform", 900 + n, b, n)
    ref = oracle.factor_batched("se99", a)
    l, e, p = ref.l.copy(), ref.e.copy(), ref.p.copy()
    want = np.array([identity_residual(a[i], l[i], e[i], p[i]) for i in range(b)])
    want = want / np.maximum(1.0, np.abs(a).max(axis=(1, 2)))
    d = m.Decomposition(torch.from_numpy(l).cuda(), torch.from_numpy(e).cuda(),
                        torch.from_numpy(p.astype(np.int64)).cuda())
    resid, flags = m.verify_batched(torch.from_numpy(a).cuda(), d)
    assert int(flags.max()) == 0
    assert np.abs(resid.cpu().numpy() - want).max() <= 1e-14, n
    rh, fh = m.verify_batched(a, m.Decomposition(l, e, p))          # HOST buffers
    assert np.array_equal(rh, resid.cpu().numpy()) and int(fh.max()) == 0
    if n < 2:
        return
    # matrix 0: p not a permutation; 1: e < 0; 2: dirty strict upper triangle; 3: NaN in L;
    # 4: one wrong entry of L (no flag, residual large); 5: out-of-range index
    p[0, 0] = p[0, 1]
    e[1, n - 1] = -1.0
    l[2, 0, n - 1] = 0.5
    l[3, n - 1, 0] = np.nan
    l[4, n - 1, 0] += 1.0
    p[5, n - 1] = n + 3
    d = m.Decomposition(torch.from_numpy(l).cuda(), torch.from_numpy(e).cuda(),
                        torch.from_numpy(p.astype(np.int64)).cuda())
    resid, flags = m.verify_batched(torch.from_numpy(a).cuda(), d)
    resid, flags = resid.cpu().numpy(), flags.cpu().numpy()
    assert flags[0] & 1 and np.isinf(resid[0])
    assert flags[1] & 2
    assert flags[2] & 4
    assert flags[3] & 8 and not resid[3] <= 1e-12
    assert flags[4] == 0 and resid[4] > 1e-3
    assert flags[5] & 1 and np.isinf(resid[5])
    assert (flags[6:] == 0).all() and (resid[6:] <= 1e-12).all()


@pytest.mark.parametrize("n", [10, 12, 128, 256])
def test_device_assembly_of_the_reference_bench_inputs(oracle, torch_cuda, n):
    """benches/bench.rs:48-53: A = Q D Q^T from the bit-exact seeded factors, assembled on the GPU
    (rank-2 reflector updates) against the reference's dense construction in numpy; then the
    factorisations of the SAME device-built bytes against the oracle (f2)."""
    torch = torch_cuda
    from modcholesky_b200 import utils
    v1, d1 = utils.bench_input_factors(n)
    dense = utils.bench_input_dense(n)
    b = 32 if n <= 128 else 8
    v, dg = utils.bench_input_batch(n, b)
    assert np.array_equal(v[0], v1) and np.array_equal(dg[0], d1)
    a = m.assemble_qdqt_batched(torch.from_numpy(v).cuda(), torch.from_numpy(dg).cuda()).cpu().numpy()
    assert np.array_equal(a, np.transpose(a, (0, 2, 1))), "assembled matrices must be exactly symmetric"
    assert np.abs(a[0] - dense).max() <= 1e-13 * np.abs(dense).max()
    ah = m.assemble_qdqt_batched(v, dg)                                # HOST buffers
    assert np.array_equal(ah, a)
    w = np.linalg.eigvalsh(a[0])
    assert np.allclose(np.sort(w), np.sort(d1), rtol=1e-9, atol=1e-9 * np.abs(d1).max())
    for algo in ALGOS:
        ref = oracle.factor_batched(algo, a)
        _assert_bit_exact(f"{algo} bench-input n={n} strict", _gpu(torch, algo, a, "strict"), ref)
        # (Q D Q^T with three reflectors is D plus a rank-6 term: 20-30 % of these matrices have a
        # pivot decision closer than 1e-9 -- rounding ties, which FAST may resolve either way)
        _assert_tolerance(f"{algo} bench-input n={n} fast", a, _gpu(torch, algo, a, "fast"), ref, min_firm=0.4)


def test_fp64_peak_probe(torch_cuda):
    dfma, dmma = m.fp64_peak_tflops()
    assert 10.0 < dfma < 80.0 and 10.0 < dmma < 80.0, (dfma, dmma)


def test_debug_bounds_build(torch_cuda):
    """compute-sanitizer is closed on the GPU pool; its stand-in is the "dbg" build
    (-DMODCHOL_DEBUG_BOUNDS=1): every shared-memory access of the kernels' accessors is checked
    against the block's window and traps otherwise.  A parity subset over every kernel class runs
    under it in a fresh process (MODCHOL_LIB selects the library at import time)."""
    import os
    import subprocess
    import sys
    from modcholesky_b200.build import CSRC, DEBUG_VARIANT
    lib = os.path.join(CSRC, f"libmodchol_b200.{DEBUG_VARIANT[0]}.so")
    if not os.path.exists(lib):
        pytest.skip("debug build not present (python -c 'import __graft_entry__ as g; g.build()')")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
import matgen, modcholesky_b200 as m, oracle
oracle.build_oracle()
assert 'dbg' in m._lib.LIB_PATH
for n in (5, 16, 24, 32, 48, 64, 100, 160, 256, 300):
    b = 8 if n <= 64 else 3
    for algo in ('gmw81', 'se90', 'se99'):
        for fam in ('uniform', 'wishart'):
            a = matgen.make(fam, 7 + n, b, n)
            ref = oracle.factor_batched(algo, a)
            ad = torch.from_numpy(a).cuda()
            print('case', algo, n, fam, flush=True)
            d = m.factor_batched(algo, ad, mode='strict')
            torch.cuda.synchronize()
            assert np.array_equal(d.l.cpu().numpy(), ref.l, equal_nan=True), (algo, n, fam)
            f = m.factor_batched(algo, ad, mode='fast')
            r, fl = m.verify_batched(ad, f)
            torch.cuda.synchronize()
            assert int(fl.max()) == 0, (algo, n, fam)
print('debug-bounds parity subset ok')
""" % (root, root)
    env = dict(os.environ, MODCHOL_LIB=lib)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "subset ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_zz_report_relaxed_bar_counts():
    """Not a check: prints how many FAST-mode comparisons of this session used one of the two
    GMW81-only relaxed bars of _assert_tolerance (run pytest with -s to see it)."""
    print(f"\n[parity] FAST-mode matrices checked: {RELAXED['checked']}, "
          f"E bar relaxed (GMW81, ||E|| < 1e-4 ||A||): {RELAXED['e_relaxed']}, "
          f"reconstruction bar relaxed (GMW81, ||A+E|| > 1e3 ||A||): {RELAXED['recon_relaxed']}")
