"""Pins the CPU oracle (oracle/modchol_oracle.c) against every literal the reference's
own in-file tests hold.  CPU only."""
import numpy as np
import pytest

from conftest import identity_residual, perm_matrix


def _case_ids(golden):
    return [c["name"] for c in golden["cases"]]


def test_reference_test_cases(golden, oracle):
    assert len(golden["cases"]) == 11
    for case in golden["cases"]:
        a = np.array(golden["matrices"][case["matrix"]]["a"], dtype=np.float64)
        r = oracle.factor(case["algo"], a)
        n = a.shape[0]
        # result-shape conventions (lib.rs:109-119): lower triangular, p a permutation
        assert np.all(np.triu(r.l, 1) == 0.0), case["name"]
        assert sorted(int(x) for x in r.p) == list(range(n)), case["name"]
        assert np.all(r.e >= 0.0), case["name"]
        assert identity_residual(a, r.l, r.e, r.p) < case["tol_identity"], case["name"]
        llt = r.l @ r.l.T
        if "u" in case:
            u = np.array(case["u"])
            assert np.abs(llt - u.T @ u).max() < case["tol_llt"], case["name"]
        if "res" in case:
            P = perm_matrix(r.p)
            res = P @ np.array(case["res"]) @ P.T
            assert np.abs(llt - res).max() < case["tol_llt"], case["name"]
        if "l" in case:
            assert np.abs(r.l - np.array(case["l"])).max() < case["tol_l"], case["name"]


# SURVEY.md section 4 "[derived] known answers" (numpy restatement made during the survey,
# independent of this C oracle): p exactly, e to 1e-13 relative, phase-2 start index.
DERIVED = [
    ("gmw81", "A3", [0, 1, 2], [2.771236166328254, 5.015611460128483, 2.2426406871192848], -1),
    ("gmw81", "A3b", [1, 0, 2], [0, 0, 3.008], -1),
    ("gmw81", "A4", [3, 0, 1, 2], [1.0333767434044603, 0.9608272410614518, 0.5563862634332963, 0], -1),
    ("gmw81", "A6", [2, 4, 1, 3, 0, 5], [0, 0, 0, 0, 0, 1.6697776494822845e-14], -1),
    ("se90", "A3", [0, 1, 2], [2, 2.2196657443588332, 2.2196657443588332], 0),
    ("se90", "A3b", [1, 0, 2], [1.5040363327267148, 0, 1.5040363327267148], 0),
    ("se90", "A4", [2, 3, 1, 0], [1049.3999999999999] * 4, 0),
    ("se90", "A6", [2, 3, 1, 0, 4, 5],
     [7.497687103356203, 7.497687103356203, 0, 0, 7.497687103356203, 7.497687103356203], 1),
    ("se99", "A3", [0, 1, 2], [2, 2.2196657443588332, 2.2196657443588332], 0),
    ("se99", "A3b", [1, 0, 2], [1.5040363327267148, 0, 1.5040363327267148], 0),
    ("se99", "A4", [3, 2, 1, 0], [0.6937638863979176, 0.6937638863979176, 0.36656864392554667, 0], 1),
    ("se99", "A6", [2, 4, 1, 3, 0, 5], [0, 0, 0, 0, 0, 0.0003139868187200542], 5),
]


@pytest.mark.parametrize("algo,name,p,e,k", DERIVED, ids=[f"{d[0]}-{d[1]}" for d in DERIVED])
def test_survey_derived_known_answers(golden, oracle, algo, name, p, e, k):
    a = np.array(golden["matrices"][name]["a"], dtype=np.float64)
    r = oracle.factor(algo, a)
    assert [int(x) for x in r.p] == p
    e = np.array(e, dtype=np.float64)
    scale = max(np.abs(e).max(), 1e-300)
    assert np.abs(r.e - e).max() <= 1e-9 * scale + 1e-24
    assert int(r.phase2_start) == k


def test_gershgorin_circles(golden, oracle):
    g = golden["gershgorin"]
    out = oracle.gershgorin_circles(np.array(g["a"]))
    assert np.abs(out - np.array(g["circles"])).max() < g["tol"]


def test_eigenvalues_2x2(oracle):
    hi, lo = oracle.eigenvalues_2x2([[2.0, 1.0], [1.0, 2.0]])
    assert (hi, lo) == (3.0, 1.0)
    hi, lo = oracle.eigenvalues_2x2([[1.0, 3.0], [3.0, 1.0]])
    assert (hi, lo) == (4.0, -2.0)


def test_edge_cases_follow_the_reference_formulas(oracle):
    """SURVEY.md section 4 '[derived] edge-case behaviour to reproduce, not fix'."""
    eps = np.finfo(np.float64).eps
    m = np.array([[-4.0]])
    r = oracle.factor("gmw81", m)
    assert r.e[0] == 8.0 and r.l[0, 0] == 2.0
    r = oracle.factor("se90", m)
    assert np.isnan(r.l[0, 0])                       # sqrt of a negative, se90.rs:84
    r = oracle.factor("se99", m)
    assert abs(r.e[0] - 4.000024221964485) < 1e-12   # se99.rs:115-119
    z = np.zeros((3, 3))
    r = oracle.factor("gmw81", z)
    assert np.all(r.e == eps) and np.allclose(r.l, np.sqrt(eps) * np.eye(3), rtol=0, atol=0)
    for algo in ("se90", "se99"):
        assert np.isnan(oracle.factor(algo, z).l).any()   # 0/0, se90.rs:86 / se99.rs:98
    for algo in ("gmw81", "se90", "se99"):
        r = oracle.factor(algo, np.eye(4))
        assert np.all(r.e == 0) and np.array_equal(r.l, np.eye(4))
        assert [int(x) for x in r.p] == [0, 1, 2, 3]
    # equal diagonals: the lowest index wins (utils.rs:57-69)
    r = oracle.factor("se99", np.diag([2.0, 5.0, 5.0, 1.0]))
    assert [int(x) for x in r.p] == [1, 2, 0, 3]


@pytest.mark.parametrize("algo", ["gmw81", "se90", "se99"])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 9, 16, 17, 32, 33, 64])
def test_properties_random(oracle, algo, n):
    """Universal acceptance identity on random symmetric indefinite and PD inputs."""
    rng = np.random.default_rng(1234 + n)
    b = 24
    a = rng.uniform(-1, 1, size=(b, n, n))
    a = np.tril(a) + np.transpose(np.tril(a, -1), (0, 2, 1))
    w = rng.standard_normal((b, n, n))
    pd = w @ np.transpose(w, (0, 2, 1)) + n * np.eye(n)
    for fam, mats in (("indef", a), ("pd", pd)):
        if n == 1 and algo == "se90" and fam == "indef":
            continue  # negative 1x1 gives NaN in SE90 (reference behaviour)
        r = oracle.factor_batched(algo, mats, threads=2)
        for i in range(b):
            if algo == "se90" and np.isnan(r.l[i]).any():
                # reference quirk (se90.rs:63-84): phase 1 has no sign test, so a trailing block
                # whose largest diagonal is negative but passes the look-ahead gets sqrt(<0).
                assert mats[i].diagonal().max() < 0 or n > 2
                continue
            assert sorted(int(x) for x in r.p[i]) == list(range(n))
            assert np.all(np.triu(r.l[i], 1) == 0.0)
            assert np.all(r.e[i] >= 0.0)
            res = identity_residual(mats[i], r.l[i], r.e[i], r.p[i])
            assert res <= 1e-12 * max(1.0, np.abs(mats[i]).max()) * n, (fam, i, res)
            if fam == "pd" and algo != "gmw81":
                assert np.all(r.e[i] == 0.0)
                assert int(r.phase2_start[i]) == -1


def test_batched_threads_agree(oracle):
    rng = np.random.default_rng(7)
    a = rng.uniform(-1, 1, size=(300, 12, 12))
    a = a + np.transpose(a, (0, 2, 1))
    r1 = oracle.factor_batched("se99", a, threads=1)
    r4 = oracle.factor_batched("se99", a, threads=4)
    assert np.array_equal(r1.l, r4.l) and np.array_equal(r1.e, r4.e) and np.array_equal(r1.p, r4.p)


def test_vectorised_generator_tables_equal_the_scalar_restatement():
    """The per-seed tables used to build batches of the reference's bench family are the same
    bits as the scalar XorShift / Uniform restatement (utils.rs:116-147)."""
    from modcholesky_b200 import utils
    hv, dg = utils.householder_vector_table(12), utils.diagonal_table(12)
    for s in (0, 2, 9, 23, 90, 255):
        assert np.array_equal(hv[s], utils.householder_vector(12, s))
        assert np.array_equal(dg[s], np.diag(utils.random_diagonal(12, (-1.0, 1000.0), 1, s)))
    dg3 = utils.diagonal_table(10, (-2.0, 5.0), 3)
    for s in (1, 77):
        assert np.array_equal(dg3[s], np.diag(utils.random_diagonal(10, (-2.0, 5.0), 3, s)))
    v, d = utils.bench_input_batch(12, 4)
    v0, d0 = utils.bench_input_factors(12)
    assert np.array_equal(v[0], v0) and np.array_equal(d[0], d0)
    m = np.diag(d0)
    for r in (2, 1, 0):
        h = np.eye(12) - 2.0 / np.dot(v0[r], v0[r]) * np.outer(v0[r], v0[r])
        m = h @ m @ h
    dense = utils.bench_input_dense(12)
    assert np.abs(m - dense).max() <= 1e-13 * np.abs(dense).max()
