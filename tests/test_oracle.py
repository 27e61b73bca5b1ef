"""CPU tests of the ORACLE itself: it is pinned before it is trusted.

No numeric log-likelihood is pinned by the reference's tests ("parity
unpinned"), so the oracle is anchored on (a) scipy / mpmath for its numerical
building blocks, (b) an independent numpy restatement of the whole path, (c) the
reference's invariants, (d) the committed golden vectors."""
import ctypes as C
import json
import math
import sys
from pathlib import Path

import numpy as np
import pytest

import fixtures
from cases import Case, compute_lh, compute_lh_root, move_root, rel_err, same_bits
from oracle_capi import MODE_ENGINE, MODE_REFERENCE, OraclePartition, gamma_cats, load_oracle

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
import np_oracle  # noqa: E402

dp = C.POINTER(C.c_double)
GOLDEN = json.loads((Path(__file__).parent / "golden" / "golden_v1.json").read_text())


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs])


def test_software_log_within_one_ulp_of_libm():
    L = load_oracle()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.random(5000), np.exp(rng.uniform(-700, 700, 5000)), [1.0, 0.5, 2.0, 5e-324, 1e-310]])
    for x in xs:
        a, b = L.rdo_log(float(x)), math.log(x)
        assert abs(a - b) <= np.spacing(abs(b)) if b != 0 else a == 0.0
    assert L.rdo_log(0.0) == -math.inf and math.isnan(L.rdo_log(-1.0)) and L.rdo_log(math.inf) == math.inf


def _q_and_p(L, r, pi, t):
    Q, E = np.zeros(16), np.zeros(16)
    L.rdo_build_q_nonrev(r.ctypes.data_as(dp), pi.ctypes.data_as(dp), Q.ctypes.data_as(dp))
    A = Q * t
    L.rdo_expm4(A.ctypes.data_as(dp), E.ctypes.data_as(dp))
    return Q.reshape(4, 4), E.reshape(4, 4)


def test_rate_matrix_and_expm_against_scipy():
    from scipy.linalg import expm
    L = load_oracle()
    rng = np.random.default_rng(1)
    for _ in range(300):
        r = rng.uniform(1e-4, 1, 12)
        pi = rng.dirichlet(np.ones(4) * 3)
        t = float(np.exp(rng.uniform(-12, 3)))
        Q, P = _q_and_p(L, r, pi, t)
        assert abs(Q.sum(1)).max() < 1e-14                      # rows sum to 0
        assert abs(-(pi * np.diag(Q)).sum() - 1) < 1e-14        # unit mean rate
        k = 0
        for i in range(4):
            for j in range(4):
                if i != j:
                    assert Q[i, j] * (-(pi * np.diag(Q * 1)).sum()) > 0
                    k += 1
        assert abs(P - expm(Q * t)).max() < 5e-14
        assert abs(P.sum(1) - 1).max() < 1e-13                  # stochastic
    _, P0 = _q_and_p(L, rng.uniform(1e-4, 1, 12), np.full(4, .25), 0.0)
    assert np.array_equal(P0, np.eye(4))                        # P(0) == I exactly (SURVEY B-15)


def test_expm_against_mpmath_50_digits():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    L = load_oracle()
    rng = np.random.default_rng(2)
    for t in (1e-6, 1e-3, 0.05, 0.7, 4.0, 30.0):
        r = rng.uniform(1e-4, 1, 12)
        pi = rng.dirichlet(np.ones(4) * 3)
        Q, P = _q_and_p(L, r, pi, t)
        ref = mp.expm(mp.matrix((Q * t).tolist()))
        ref = np.array([[float(ref[i, j]) for j in range(4)] for i in range(4)])
        assert abs(P - ref).max() < 2e-14


def test_gamma_categories_against_scipy():
    from scipy import special, stats
    for alpha in (0.2, 0.5, 1.0, 2.0, 10.0):
        for K in (2, 4, 8):
            mean = gamma_cats(alpha, K, 0)
            qs = stats.gamma.ppf(np.arange(1, K) / K, alpha, scale=1 / alpha)
            cdf = np.concatenate([[0], special.gammainc(alpha + 1, qs * alpha), [1]])
            assert abs(mean - np.diff(cdf) * K).max() < 1e-6
            assert abs(mean.mean() - 1) < 1e-9
            med = gamma_cats(alpha, K, 1)
            ref = stats.gamma.ppf((2 * np.arange(K) + 1) / (2 * K), alpha, scale=1 / alpha)
            assert abs(med - ref / ref.mean()).max() < 1e-6
    assert np.array_equal(gamma_cats(0.7, 1, 0), [1.0])


def test_golden_gamma_and_expm_vectors():
    L = load_oracle()
    for g in GOLDEN["gamma_cats"]:
        assert same_bits(gamma_cats(g["alpha"], g["k"], g["mode"]), unhex(g["rates"]))
    for g in GOLDEN["expm"]:
        Q, P = _q_and_p(L, unhex(g["rates"]), unhex(g["freqs"]), float.fromhex(g["t"]))
        assert same_bits(Q.ravel(), unhex(g["Q"])) and same_bits(P.ravel(), unhex(g["P"]))


def test_pairwise_sum_is_the_canonical_tree():
    L = load_oracle()
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 7, 8, 9, 1000, 1025):
        v = rng.normal(size=n) * 1e3
        N = 1
        while N < n:
            N *= 2
        w = np.concatenate([v, np.zeros(N - n)])
        while len(w) > 1:
            w = w[0::2] + w[1::2]
        assert L.rdo_pairwise_sum(v.ctypes.data_as(dp), n) == w[0]
    assert L.rdo_pairwise_sum(None, 0) == 0.0


@pytest.mark.parametrize("n,S,K,data", [(5, 7, 4, "evolved"), (10, 400, 4, "evolved"), (33, 300, 2, "ambiguous"),
                                        (300, 64, 4, "iid"), (12, 50, 1, "evolved")])
def test_oracle_against_independent_numpy_restatement(n, S, K, data):
    c = Case(n, S, K, seed=5, data=data, weights="random")
    o = OraclePartition(n, S, K)
    c.setup(o)
    sched = c.full_schedule(1, 0.3)
    l_ref, ps = compute_lh(o, sched, c.root_clv, c.root_scaler, persite=True, mode=MODE_REFERENCE)
    l_eng = o.root_loglikelihood(c.root_clv, c.root_scaler, mode=MODE_ENGINE)
    ops, pm, br = sched
    tips = {c.tree.tip_index(l): s for l, s in c.aln.items()}
    l_np, ps_np = np_oracle.loglikelihood(n, tips, [op.astuple() for op in ops], pm, br, c.rates, c.freqs,
                                          c.cat_rates, c.cat_weights, c.weights, c.root_clv, c.root_scaler)
    assert abs(l_ref - l_np) <= 1e-12 * abs(l_np)
    assert rel_err(ps, ps_np) <= 1e-12
    assert abs(l_ref - l_eng) <= 1e-12 * abs(l_ref)   # the two arithmetic modes agree far inside 1e-9
    if data == "iid":
        assert o.get_scaler(c.root_scaler).max() >= 1  # underflow rescaling exercised


@pytest.mark.parametrize("n,S,K,data", [(6, 24, 4, "evolved"), (9, 16, 2, "ambiguous"), (40, 8, 4, "iid")])
def test_oracle_against_exact_arithmetic(n, S, K, data):
    """the whole chain -- rate matrix, matrix exponential, pruning over the schedule, weighted root
    log-likelihood -- re-evaluated in 60-digit arithmetic (mpmath; no rescaling needed there): the
    oracle's fp64 result, in both arithmetic modes, is within 1e-12 relative of the exact value, i.e.
    three orders inside the north-star tolerance"""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60
    c = Case(n, S, K, seed=17, data=data, weights="random")
    o = OraclePartition(n, S, K)
    c.setup(o)
    sched = c.full_schedule(2 % c.tree.root_count, 0.35)
    l_ref, ps = compute_lh(o, sched, c.root_clv, c.root_scaler, persite=True, mode=MODE_REFERENCE)
    l_eng = o.root_loglikelihood(c.root_clv, c.root_scaler, mode=MODE_ENGINE)
    ops, pm, br = sched
    r, f = [mp.mpf(float(x)) for x in c.rates], [mp.mpf(float(x)) for x in c.freqs]
    Q = mp.zeros(4, 4)
    k = 0
    for i in range(4):
        for j in range(4):
            if i != j:
                Q[i, j] = r[k] * f[j]
                k += 1
    for i in range(4):
        Q[i, i] = -sum(Q[i, j] for j in range(4) if j != i)
    mu = -sum(f[i] * Q[i, i] for i in range(4))
    Q = Q / mu
    P = {int(i): [mp.expm(Q * (mp.mpf(float(cr)) * mp.mpf(float(t)))) for cr in c.cat_rates] for i, t in zip(pm, br)}
    clv = {}
    for label, seq in c.aln.items():
        m = [np_oracle.NT[chr(ch).upper()] for ch in seq]
        clv[c.tree.tip_index(label)] = [[[mp.mpf((mm >> i) & 1) for i in range(4)] for _ in range(K)] for mm in m]
    for op in ops:
        p, _, c1, m1, _, c2, m2, _ = op.astuple()
        out = []
        for s in range(S):
            row = []
            for kk in range(K):
                x = [sum(P[m1][kk][i, j] * clv[c1][s][kk][j] for j in range(4)) for i in range(4)]
                y = [sum(P[m2][kk][i, j] * clv[c2][s][kk][j] for j in range(4)) for i in range(4)]
                row.append([x[i] * y[i] for i in range(4)])
            out.append(row)
        clv[p] = out
    persite = []
    for s in range(S):
        term = sum(mp.mpf(float(c.cat_weights[kk])) * sum(f[i] * clv[c.root_clv][s][kk][i] for i in range(4))
                   for kk in range(K))
        persite.append(mp.log(term) * int(c.weights[s]))
    exact = sum(persite)
    assert abs(mp.mpf(l_ref) - exact) <= mp.mpf("1e-12") * abs(exact)
    assert abs(mp.mpf(l_eng) - exact) <= mp.mpf("1e-12") * abs(exact)
    assert max(abs(mp.mpf(float(a)) - b) / abs(b) for a, b in zip(ps, persite) if b != 0) <= mp.mpf("1e-12")


def test_golden_fixture_loglikelihoods():
    """the reference's bundled fixtures (test/data): the oracle reproduces the committed vectors bit for bit"""
    for key, rec in GOLDEN["fixtures"].items():
        name, K = key.split(":K")
        case = fixtures.FixtureCase(fixtures.load(name), int(K))
        assert case.S == rec["patterns"] and case.n == rec["taxa"] and case.tree.root_count == rec["roots"]
        o = OraclePartition(case.n, case.S, int(K))
        case.setup(o)
        for g in rec["lh"][:12]:
            a = compute_lh(o, case.full_schedule(g["root"], g["ratio"]), case.root_clv, case.root_scaler,
                           mode=MODE_REFERENCE)
            b = o.root_loglikelihood(case.root_clv, case.root_scaler, mode=MODE_ENGINE)
            assert a.hex() == g["reference"] and b.hex() == g["engine"]


def test_golden_synthetic_cases():
    for g in GOLDEN["synthetic"]:
        case = Case(g["n"], g["S"], g["K"], seed=g["seed"], data=g["data"], weights=g["weights"])
        o = OraclePartition(case.n, case.S, case.K)
        case.setup(o)
        a, ps = compute_lh(o, case.full_schedule(1, 0.3), case.root_clv, case.root_scaler, persite=True,
                           mode=MODE_REFERENCE)
        b = o.root_loglikelihood(case.root_clv, case.root_scaler, mode=MODE_ENGINE)
        c = compute_lh_root(o, case.derivative_schedule(1, 0.7), case.root_clv, case.root_scaler, mode=MODE_ENGINE)
        assert a.hex() == g["reference"] and b.hex() == g["engine"] and c.hex() == g["engine_ratio_0.7"]
        assert same_bits(ps[:8], unhex(g["persite_head"]))
        assert int(o.get_scaler(case.root_scaler).max()) == g["max_scaler"]


def test_reference_invariants_hold_for_the_oracle():
    """test/src/model.cpp:59-75 (finite, < 0, bit-reproducible), :271-288 (full == root-only)"""
    case = fixtures.FixtureCase(fixtures.load("10.fasta"), 1)
    o = OraclePartition(case.n, case.S, 1)
    case.setup(o)
    for rid in range(case.tree.root_count):
        a = compute_lh(o, case.full_schedule(rid), case.root_clv, case.root_scaler)
        b = compute_lh(o, case.full_schedule(rid), case.root_clv, case.root_scaler)
        c = compute_lh_root(o, case.derivative_schedule(rid, 0.5), case.root_clv, case.root_scaler)
        assert math.isfinite(a) and a < 0 and a == b and a == c


def test_root_invariance_under_reversible_parameters():
    """test/src/model.cpp:367-387: all rates 1 + uniform pi (JC) => every root has the same logL (rel 1.19e-5)"""
    case = fixtures.FixtureCase(fixtures.load("101.phy"), 1)
    case.rates = np.ones(12)
    o = OraclePartition(case.n, case.S, 1)
    case.setup(o)
    compute_lh(o, case.full_schedule(0), case.root_clv, case.root_scaler)
    vals = []
    for rid in range(0, case.tree.root_count, 9):
        move_root(o, case.move_schedule(rid))
        vals.append(compute_lh(o, case.full_schedule(rid), case.root_clv, case.root_scaler))
    assert np.ptp(vals) <= 1.19e-5 * abs(vals[0])


@pytest.mark.parametrize("flags", range(8))
def test_reference_invariants_hold_under_every_q_convention(flags):
    """the pi-multiplication, the slot order of the 12 rates and the normalisation of the
    non-reversible Q are unverified readings of coraxlib -- run-time switches of the oracle
    (rdo_set_q_convention).  Every invariant the reference's own tests hold (test/src/model.cpp:59-75
    finite / negative / reproducible, :271-288 full traversal == root-only evaluation, :367-387 root
    invariance under JC parameters) holds under each of the 8 variants, so none of them can be ruled
    out from the reference's tests alone; the default is the one the CUDA engine implements."""
    from oracle_capi import load_oracle
    L = load_oracle()
    assert L.rdo_get_q_convention() == 0
    L.rdo_set_q_convention(flags)
    try:
        case = fixtures.FixtureCase(fixtures.load("10.fasta"), 4)
        case.freqs = np.array([0.1, 0.2, 0.3, 0.4])  # (under uniform pi the first switch is a no-op)
        o = OraclePartition(case.n, case.S, 4)
        case.setup(o)
        vals = []
        for rid in range(0, case.tree.root_count, 4):
            a = compute_lh(o, case.full_schedule(rid), case.root_clv, case.root_scaler)
            b = compute_lh(o, case.full_schedule(rid), case.root_clv, case.root_scaler)
            c = compute_lh_root(o, case.derivative_schedule(rid, 0.5), case.root_clv, case.root_scaler)
            assert math.isfinite(a) and a < 0 and a == b and a == c
            vals.append(a)
        if flags:
            # a different convention is a different model: the variants are distinguishable by value
            L.rdo_set_q_convention(0)
            base = compute_lh(o, case.full_schedule(0), case.root_clv, case.root_scaler)
            L.rdo_set_q_convention(flags)
            assert abs(base - vals[0]) > 1e-6 * abs(base)
        jc = fixtures.FixtureCase(fixtures.load("10.fasta"), 1)
        jc.rates = np.ones(12)
        o1 = OraclePartition(jc.n, jc.S, 1)
        jc.setup(o1)
        same = [compute_lh(o1, jc.full_schedule(rid), jc.root_clv, jc.root_scaler) for rid in range(0, jc.tree.root_count, 3)]
        assert np.ptp(same) <= 1.19e-5 * abs(same[0])
    finally:
        L.rdo_set_q_convention(0)


def test_empty_and_tiny_partitions():
    o = OraclePartition(4, 0, 4)
    assert o.root_loglikelihood(6, 2) == 0.0
    case = fixtures.FixtureCase(fixtures.load("single"), 4)
    assert case.S == 1 and case.n == 4
    o = OraclePartition(4, 1, 4)
    case.setup(o)
    v = compute_lh(o, case.full_schedule(2, 0.5), case.root_clv, case.root_scaler)
    assert math.isfinite(v) and v < 0


def test_illegal_state_is_rejected():
    o = OraclePartition(4, 5, 1)
    with pytest.raises(ValueError):
        o.set_tip_states(0, b"ACG!T")
