"""GPU parity tests: the CUDA engine (through the C ABI of include/rdk.h) against
the CPU oracle on the same seeded inputs.

Bars (BASELINE.md section 5 / north_star):
  * fp64 log-likelihoods, per site and total: <= 1e-9 relative against the
    oracle in REFERENCE arithmetic (libm log, serial site sum);
  * stronger, by construction: bit-for-bit equality of P-matrices, CLVs,
    scalers, per-site and total log-likelihood against the oracle in ENGINE
    arithmetic (spec'd software log + canonical pairwise tree), which is what
    makes the chosen root / LWR ranking / optimised alpha identical.
"""
import numpy as np
import pytest

import cases
from cases import Case, bits, compute_lh, compute_lh_root, move_root, rel_err, same_bits
from oracle_capi import MODE_ENGINE, MODE_REFERENCE, OraclePartition

pytestmark = pytest.mark.gpu

RTOL = 1e-9  # north_star tolerance for fp64 log-likelihoods


def make(case):
    from root_digger_b200.capi import Partition
    g = Partition(case.n, case.S, case.K)
    o = OraclePartition(case.n, case.S, case.K)
    case.setup(g)
    case.setup(o)
    return g, o


CASES = [
    # n_taxa, sites, K, data, weights
    (4, 1, 1, "evolved", "ones"),
    (5, 7, 4, "evolved", "random"),
    (10, 1000, 1, "evolved", "ones"),
    (10, 991, 4, "evolved", "random"),
    (33, 2049, 2, "ambiguous", "random"),
    (64, 4097, 8, "evolved", "ones"),
    (101, 1630, 4, "ambiguous", "random"),
    (300, 523, 4, "iid", "ones"),      # deep underflow: scalers fire
    (40, 300, 16, "evolved", "ones"),
    (24, 100, 32, "evolved", "ones"),
    # any --rate-cats (reference src/model.cpp:159-168 passes K straight to corax_partition_create):
    # the device pads to the next divisor of 32 with zero-weight copies of category 0
    (30, 700, 3, "evolved", "random"),
    (300, 257, 3, "iid", "ones"),      # padding must not disturb the all-entries-small rescaling test
    (20, 400, 5, "ambiguous", "random"),
    (16, 300, 6, "evolved", "ones"),
    (12, 200, 11, "evolved", "ones"),
]


@pytest.mark.parametrize("n,S,K,data,weights", CASES)
def test_full_evaluation_matches_oracle(n, S, K, data, weights):
    case = Case(n, S, K, seed=1000 + n + S, data=data, weights=weights)
    g, o = make(case)
    sched = case.full_schedule(0, 0.5)
    lg, pg = compute_lh(g, sched, case.root_clv, case.root_scaler, persite=True)
    lo_ref, po_ref = compute_lh(o, sched, case.root_clv, case.root_scaler, persite=True, mode=MODE_REFERENCE)
    lo_eng, po_eng = o.root_loglikelihood(case.root_clv, case.root_scaler, persite=True, mode=MODE_ENGINE)
    assert np.isfinite(lg) and lg < 0
    # contractual tolerance
    assert abs(lg - lo_ref) <= RTOL * abs(lo_ref)
    assert rel_err(pg, po_ref) <= RTOL
    # bit-for-bit against the oracle in engine arithmetic
    ops, pm, br = sched
    for mi in pm:
        assert same_bits(g.get_pmatrix(int(mi)), o.get_pmatrix(int(mi))), f"P-matrix {mi}"
    for op in ops:
        assert same_bits(g.get_clv(op.parent_clv_index), o.get_clv(op.parent_clv_index)), "CLV"
        assert np.array_equal(g.get_scaler(op.parent_scaler_index), o.get_scaler(op.parent_scaler_index))
    assert same_bits(pg, po_eng)
    assert same_bits([lg], [lo_eng])
    if data == "iid" and n >= 300:
        assert g.get_scaler(case.root_scaler).max() >= 1, "the underflow case must exercise the scalers"
    # reference invariant: bit-identical on a second call (test/src/model.cpp:59-75)
    lg2 = compute_lh(g, sched, case.root_clv, case.root_scaler)
    assert same_bits([lg], [lg2])


def test_tip_clv_readback_and_frequencies():
    case = Case(12, 333, 4, seed=7, data="ambiguous", weights="random")
    g, o = make(case)
    for t in range(case.n):
        assert same_bits(g.get_clv(t), o.get_clv(t))
    fg, fo = g.empirical_frequencies(), o.empirical_frequencies()
    assert abs(fg.sum() - 1) < 1e-12
    assert np.allclose(fg, fo, rtol=1e-13, atol=0)


@pytest.mark.parametrize("n,S,K", [(10, 991, 4), (50, 3001, 4), (17, 64, 1)])
def test_root_only_and_full_paths_agree(n, S, K):
    """test/src/model.cpp:271-288: compute_lh == compute_lh_root (zero tolerance)"""
    case = Case(n, S, K, seed=n * S)
    g, o = make(case)
    for rid in range(case.tree.root_count):
        full = compute_lh(g, case.full_schedule(rid, 0.5), case.root_clv, case.root_scaler)
        root = compute_lh_root(g, case.derivative_schedule(rid, 0.5), case.root_clv, case.root_scaler)
        assert same_bits([full], [root])
        if rid % 7 == 0:
            oo = compute_lh(o, case.full_schedule(rid, 0.5), case.root_clv, case.root_scaler, mode=MODE_ENGINE)
            assert same_bits([full], [oo])


@pytest.mark.parametrize("n,S,K,data", [(10, 991, 4, "evolved"), (60, 2000, 4, "ambiguous"), (300, 257, 4, "iid")])
def test_move_root_sequence_matches_oracle(n, S, K, data):
    """suggest_roots_lh (src/model.cpp:865-889): move_root + compute_lh_root per root"""
    case = Case(n, S, K, seed=31 * n, data=data)
    g, o = make(case)
    s0 = case.full_schedule(0, 0.5)
    compute_lh(g, s0, case.root_clv, case.root_scaler)
    compute_lh(o, s0, case.root_clv, case.root_scaler)
    rng = np.random.default_rng(5)
    order = list(range(case.tree.root_count))
    rng.shuffle(order)
    for rid in order[:40]:
        ms = case.move_schedule(rid, 0.5)
        ds = case.derivative_schedule(rid, 0.37)
        move_root(g, ms)
        move_root(o, ms)
        a = compute_lh_root(g, ds, case.root_clv, case.root_scaler)
        b_ref = compute_lh_root(o, ds, case.root_clv, case.root_scaler, mode=MODE_REFERENCE)
        b_eng = o.root_loglikelihood(case.root_clv, case.root_scaler, mode=MODE_ENGINE)
        assert abs(a - b_ref) <= RTOL * abs(b_ref)
        assert same_bits([a], [b_eng])


def test_root_invariance_under_reversible_model():
    """test/src/model.cpp:367-387: all-ones rates + uniform pi => every root gives the same logL"""
    case = Case(30, 800, 4, seed=99, data="ambiguous")
    case.rates = np.ones(12)
    case.freqs = np.full(4, 0.25)
    g, o = make(case)
    vals = [compute_lh(g, case.full_schedule(r, 0.5), case.root_clv, case.root_scaler)
            for r in range(case.tree.root_count)]
    assert np.ptp(vals) <= 1.2e-5 * abs(vals[0])


def test_multi_candidate_root_evaluation():
    """rdk_root_loglikelihood_multi == repeated compute_lh_root, state untouched"""
    case = Case(40, 1500, 4, seed=3)
    g, o = make(case)
    s0 = case.full_schedule(5, 0.5)
    compute_lh(g, s0, case.root_clv, case.root_scaler)
    compute_lh(o, s0, case.root_clv, case.root_scaler)
    op, pm, br = case.derivative_schedule(5, 0.5)
    total = br.sum()
    ratios = np.array([0.0, 1e-8, 0.25, 0.5, 0.5 + 1e-8, 0.75, 1.0 - 1e-8, 1.0])
    pairs = np.stack([total * ratios, total * (1 - ratios)], axis=1)
    before = g.get_clv(op.parent_clv_index).copy()
    got = g.root_loglikelihood_multi(op, pairs)
    want = o.root_loglikelihood_multi(op, pairs, mode=MODE_ENGINE)
    assert same_bits(got, want)
    assert same_bits(before, g.get_clv(op.parent_clv_index))
    # zero-length root branch: P(0) must be exactly the identity (SURVEY B-15)
    g.update_prob_matrices([int(pm[0])], [0.0])
    assert np.array_equal(g.get_pmatrix(int(pm[0])), np.tile(np.eye(4), (case.K, 1, 1)))


@pytest.mark.parametrize("n,S,K,data", [(12, 500, 4, "evolved"), (80, 1200, 4, "ambiguous"), (300, 129, 2, "iid")])
def test_sweep_equals_sequential_calls(n, S, K, data):
    case = Case(n, S, K, seed=17 + n, data=data)
    g, o = make(case)
    s0 = case.full_schedule(0, 0.5)
    compute_lh(g, s0, case.root_clv, case.root_scaler)
    compute_lh(o, s0, case.root_clv, case.root_scaler)
    roots = list(range(case.tree.root_count))
    sw = case.sweep_schedule(roots, 0.5)
    got = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler)
    want = o.sweep_root_placements(*sw, case.root_clv, case.root_scaler, mode=MODE_ENGINE)
    want_ref = o.sweep_root_placements(*case.sweep_schedule(roots, 0.5), case.root_clv, case.root_scaler,
                                       mode=MODE_REFERENCE)
    assert same_bits(got, want)
    assert rel_err(got, want_ref) <= RTOL
    # ranking of candidate roots identical
    assert np.array_equal(np.argsort(-got, kind="stable"), np.argsort(-want, kind="stable"))
    # state afterwards = state after the last placement
    for idx in (case.root_clv,):
        assert same_bits(g.get_clv(idx), o.get_clv(idx))


@pytest.mark.parametrize("n,S,K,data", [(12, 500, 4, "evolved"), (80, 1200, 4, "ambiguous"), (300, 129, 2, "iid"),
                                        (150, 700, 4, "iid")])
def test_directed_sweep_equals_the_reference_loop(n, S, K, data):
    """the directed-CLV sweep (rooted_tree_t::generate_sweep_operations + RDK_SWEEP_KEEP_ROOT) on the
    engine == the reference's move_root + compute_lh_root loop on the oracle, bit for bit, from a
    non-trivial current root; scalers fire in the iid cases; partition state is left untouched"""
    from root_digger_b200.capi import RDK_SWEEP_KEEP_ROOT, Partition
    case = Case(n, S, K, seed=23 + n, data=data, weights="random")
    lay = case.tree.sweep_layout()
    kw = dict(clv_buffers=lay["clv_buffers"], scale_buffers=lay["scale_buffers"], prob_matrices=lay["prob_matrices"])
    g, o = Partition(case.n, case.S, case.K, **kw), OraclePartition(case.n, case.S, case.K, **kw)
    case.setup(g)
    case.setup(o)
    start = case.tree.root_count // 3
    s0 = case.full_schedule(start, 0.4)
    lh0 = compute_lh(g, s0, case.root_clv, case.root_scaler)
    compute_lh(o, s0, case.root_clv, case.root_scaler)
    *sw, pos = case.tree.generate_sweep_operations(layout=lay)
    assert sorted(pos.tolist()) == list(range(case.tree.root_count))
    # ~1 CLV operation + 1 root evaluation per placement
    assert len(sw[4]) <= 2 * case.tree.root_count
    root_before = g.get_clv(case.root_clv).copy()
    got = np.empty(len(pos))
    got[pos] = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler, flags=RDK_SWEEP_KEEP_ROOT)
    same = np.empty(len(pos))
    same[pos] = o.sweep_root_placements(*sw, case.root_clv, case.root_scaler, mode=MODE_ENGINE)
    assert same_bits(got, same)                       # same schedule on the oracle
    assert same_bits(root_before, g.get_clv(case.root_clv))
    assert case.tree.rooted and compute_lh_root(g, case.derivative_schedule(start, 0.4), case.root_clv,
                                                case.root_scaler) == lh0
    # the reference's loop on the oracle, in root-id order
    compute_lh(o, case.full_schedule(start, 0.4), case.root_clv, case.root_scaler)
    roots = list(range(case.tree.root_count))
    want = o.sweep_root_placements(*case.sweep_schedule(roots, 0.5), case.root_clv, case.root_scaler, mode=MODE_ENGINE)
    assert same_bits(got, want)
    compute_lh(o, case.full_schedule(start, 0.4), case.root_clv, case.root_scaler)
    want_ref = o.sweep_root_placements(*case.sweep_schedule(roots, 0.5), case.root_clv, case.root_scaler,
                                       mode=MODE_REFERENCE)
    assert rel_err(got, want_ref) <= RTOL
    if data == "iid":
        assert max(int(g.get_scaler(lay["scaler0"]).max()), int(g.get_scaler(case.root_scaler).max())) > 0


@pytest.mark.parametrize("n,S,K,data,chunks", [(40, 700, 4, "evolved", 2), (150, 700, 4, "iid", 5),
                                               (300, 129, 2, "iid", 16), (64, 3000, 4, "ambiguous", 3)])
def test_chunked_sweep_equals_the_unchunked_sweep(n, S, K, data, chunks):
    """rdk_sweep_root_placements_chunks: the directed sweep cut into independent chunks of placements,
    walked side by side in ONE launch (gridDim.y = chunks), returns the bits of the unchunked directed
    sweep -- hence of the reference's move_root + compute_lh_root loop -- and leaves the partition's own
    CLVs untouched; the launch count shows the chunks really shared a launch"""
    from root_digger_b200.capi import RDK_SWEEP_KEEP_ROOT, Partition
    case = Case(n, S, K, seed=5 + n, data=data, weights="random")
    lay = case.tree.sweep_layout(chunks)
    kw = dict(clv_buffers=lay["clv_buffers"], scale_buffers=lay["scale_buffers"], prob_matrices=lay["prob_matrices"])
    g, o = Partition(case.n, case.S, case.K, **kw), OraclePartition(case.n, case.S, case.K, **kw)
    case.setup(g)
    case.setup(o)
    start = case.tree.root_count // 2
    s0 = case.full_schedule(start, 0.3)
    lh0 = compute_lh(g, s0, case.root_clv, case.root_scaler)
    compute_lh(o, s0, case.root_clv, case.root_scaler)
    *sw, pos = case.tree.generate_sweep_operations(layout=lay)
    want = np.empty(len(pos))
    want[pos] = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler, flags=RDK_SWEEP_KEEP_ROOT)
    *csw, cpos, coff = case.tree.generate_chunked_sweep_operations(layout=lay)
    assert len(coff) - 1 == min(chunks, case.tree.root_count // 8) and sorted(cpos.tolist()) == sorted(pos.tolist())
    before = g.stats()["program_launches"]
    root_before = g.get_clv(case.root_clv).copy()
    got = np.empty(len(cpos))
    got[cpos] = g.sweep_root_placements(*csw, case.root_clv, case.root_scaler, flags=RDK_SWEEP_KEEP_ROOT,
                                        chunk_offsets=coff)
    assert g.stats()["program_launches"] == before + 1
    assert same_bits(got, want)
    assert same_bits(root_before, g.get_clv(case.root_clv))
    assert compute_lh_root(g, case.derivative_schedule(start, 0.3), case.root_clv, case.root_scaler) == lh0
    # the same chunked schedule run in order on the oracle, and the reference's loop
    ora = np.empty(len(cpos))
    ora[cpos] = o.sweep_root_placements(*csw, case.root_clv, case.root_scaler, mode=MODE_ENGINE)
    assert same_bits(got, ora)
    # without RDK_SWEEP_KEEP_ROOT the chunks share the root buffers: the engine runs them in order
    got2 = np.empty(len(cpos))
    got2[cpos] = g.sweep_root_placements(*csw, case.root_clv, case.root_scaler, flags=0, chunk_offsets=coff)
    assert same_bits(got2, want)


@pytest.mark.parametrize("n,S,K,data", [(60, 900, 4, "evolved"), (150, 700, 4, "iid"), (33, 2500, 8, "ambiguous")])
def test_discarded_sweep_buffers_do_not_change_the_values(n, S, K, data, monkeypatch):
    """RDK_SWEEP_DISCARD lets the engine keep a directed CLV in registers instead of storing it when no
    later operation of the sweep reads it back (csrc/rdk_lower.hpp liveness): same bits as the sweep
    that stores everything, in one launch and -- RDK_SWEEP_MAX_SLOTS forces it -- cut into batches, where
    a buffer is only scratch for a batch if no LATER batch reads it; stores are really dropped"""
    from root_digger_b200.capi import RDK_SWEEP_DISCARD, RDK_SWEEP_KEEP_ROOT, Partition
    case = Case(n, S, K, seed=91 + n, data=data, weights="random")
    lay = case.tree.sweep_layout()
    kw = dict(clv_buffers=lay["clv_buffers"], scale_buffers=lay["scale_buffers"], prob_matrices=lay["prob_matrices"])
    g = Partition(case.n, case.S, case.K, **kw)
    case.setup(g)
    s0 = case.full_schedule(2, 0.6)
    lh0 = compute_lh(g, s0, case.root_clv, case.root_scaler)
    *sw, pos = case.tree.generate_sweep_operations(layout=lay)
    want = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler, flags=RDK_SWEEP_KEEP_ROOT)
    g.reset_stats()
    got = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler, flags=RDK_SWEEP_KEEP_ROOT | RDK_SWEEP_DISCARD)
    st = g.stats()
    assert same_bits(got, want)
    assert st["stores_elided"] > case.n // 2 and st["program_launches"] == 1
    for slots in (7, 64):
        monkeypatch.setenv("RDK_SWEEP_MAX_SLOTS", str(slots))
        g.reset_stats()
        got = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler, flags=RDK_SWEEP_KEEP_ROOT | RDK_SWEEP_DISCARD)
        st = g.stats()
        assert same_bits(got, want), slots
        assert st["program_launches"] == -(-len(pos) // slots) and st["stores_elided"] > 0
        monkeypatch.delenv("RDK_SWEEP_MAX_SLOTS")
    # the partition's own state is what it was: the root-only evaluation of the current root still agrees
    assert compute_lh_root(g, case.derivative_schedule(2, 0.6), case.root_clv, case.root_scaler) == lh0


@pytest.mark.parametrize("n,S,K,data", [(64, 2000, 4, "evolved"), (300, 523, 4, "iid"), (40, 900, 3, "ambiguous")])
def test_lazily_materialised_evaluations_are_eager_semantics(n, S, K, data):
    """rdk_partition_set_lazy (default on): a full traversal that is only asked for its root
    log-likelihood -- compute_lh_partition inside the BFGS closures, reference src/model.cpp:455-476 --
    keeps most CLVs in registers; whatever is called next sees exactly the state eager execution leaves:
    repeated evaluations with new parameters (the kept program is superseded: no replay), root-only
    evaluation, root move, CLV / scaler read-back, the placement sweep (one replay each), all bit for
    bit against an eager partition and the oracle"""
    from root_digger_b200.capi import RDK_SWEEP_DISCARD, RDK_SWEEP_KEEP_ROOT, Partition
    case = Case(n, S, K, seed=300 + n, data=data, weights="random")
    lay = case.tree.sweep_layout()
    kw = dict(clv_buffers=lay["clv_buffers"], scale_buffers=lay["scale_buffers"], prob_matrices=lay["prob_matrices"])
    lazy, eager = Partition(case.n, case.S, case.K, **kw), Partition(case.n, case.S, case.K, **kw)
    o = OraclePartition(case.n, case.S, case.K, **kw)
    for part in (lazy, eager, o):
        case.setup(part)
    eager.set_lazy(False)
    sched = case.full_schedule(4, 0.3)
    rng = np.random.default_rng(5)
    lazy.reset_stats()
    # a BFGS-like run: same traversal, new substitution rates every time
    for it in range(5):
        rates = case.rates * (1.0 + 0.1 * rng.random(12))
        for part in (lazy, eager, o):
            part.set_subst_params(rates)
        a = compute_lh(lazy, sched, case.root_clv, case.root_scaler)
        b = compute_lh(eager, sched, case.root_clv, case.root_scaler)
        c = compute_lh(o, sched, case.root_clv, case.root_scaler, mode=MODE_ENGINE)
        assert same_bits([a], [b]) and same_bits([a], [c]), it
    st = lazy.stats()
    # the first traversal of a streak is eager (a lone compute_lh would only have to be replayed)
    assert st["lazy_evaluations"] == 4 and st["materializations"] == 0 and st["stores_elided"] >= 4 * (n // 4)
    assert eager.stats()["lazy_evaluations"] == 0
    # root-only evaluation at another position of the same branch: reads the two root children
    ds = case.derivative_schedule(4, 0.7)
    assert same_bits([compute_lh_root(lazy, ds, case.root_clv, case.root_scaler)],
                     [compute_lh_root(eager, ds, case.root_clv, case.root_scaler)])
    assert lazy.stats()["materializations"] == 1
    # every CLV and scale buffer is what eager execution left (and what the oracle has)
    for _ in range(2):  # the second one is lazy
        compute_lh(lazy, sched, case.root_clv, case.root_scaler)
    compute_lh(eager, sched, case.root_clv, case.root_scaler)
    compute_lh(o, sched, case.root_clv, case.root_scaler)
    for op in sched[0]:
        assert same_bits(lazy.get_clv(op.parent_clv_index), o.get_clv(op.parent_clv_index))
        assert np.array_equal(lazy.get_scaler(op.parent_scaler_index), o.get_scaler(op.parent_scaler_index))
    assert lazy.stats()["materializations"] == 2
    # lazy evaluation, then a root move + root-only evaluation elsewhere
    for _ in range(2):
        compute_lh(lazy, sched, case.root_clv, case.root_scaler)
    assert lazy.stats()["lazy_evaluations"] == 6
    ms = case.move_schedule(9, 0.5)
    for part in (lazy, eager):
        move_root(part, ms)
    ds = case.derivative_schedule(9, 0.25)
    assert same_bits([compute_lh_root(lazy, ds, case.root_clv, case.root_scaler)],
                     [compute_lh_root(eager, ds, case.root_clv, case.root_scaler)])
    # lazy evaluation, then the directed sweep (reads every CLV of the partition)
    s1 = case.full_schedule(1, 0.5)
    for part in (lazy, lazy, eager):
        compute_lh(part, s1, case.root_clv, case.root_scaler)
    *sw, pos = case.tree.generate_sweep_operations(layout=lay)
    fl = RDK_SWEEP_KEEP_ROOT | RDK_SWEEP_DISCARD
    assert same_bits(lazy.sweep_root_placements(*sw, case.root_clv, case.root_scaler, flags=fl),
                     eager.sweep_root_placements(*sw, case.root_clv, case.root_scaler, flags=fl))
    # per-site output requested: never lazy
    before = lazy.stats()["lazy_evaluations"]
    _, ps = compute_lh(lazy, s1, case.root_clv, case.root_scaler, persite=True)
    _, pe = compute_lh(eager, s1, case.root_clv, case.root_scaler, persite=True)
    assert same_bits(ps, pe) and lazy.stats()["lazy_evaluations"] == before


@pytest.mark.parametrize("n,S,K,data", [(64, 2000, 4, "evolved"), (300, 523, 4, "iid"), (130, 900, 3, "ambiguous"),
                                        (500, 12500, 4, "evolved")])
def test_subtree_groups_are_array_order_semantics(n, S, K, data):
    """rdk_partition_set_subtree_groups: disjoint subtrees of a traversal walked side by side (one
    blockIdx.y each) and then the operations that join them -- every CLV, scale buffer, per-site and
    total log-likelihood has the bits of the operations executed in array order (corax_update_clvs,
    reference src/model.cpp:402) and of the oracle, eager and inside a lazily materialised streak;
    the default (the engine's cost model) groups the long traversals of these small shards by itself"""
    from root_digger_b200.capi import Partition
    case = Case(n, S, K, seed=700 + n, data=data, weights="random")
    o = OraclePartition(case.n, case.S, case.K)
    case.setup(o)
    sched = case.full_schedule(5, 0.3)
    want, want_ps = compute_lh(o, sched, case.root_clv, case.root_scaler, mode=MODE_ENGINE, persite=True)
    ds = case.derivative_schedule(5, 0.7)
    want_root = compute_lh_root(o, ds, case.root_clv, case.root_scaler, mode=MODE_ENGINE)
    compute_lh(o, sched, case.root_clv, case.root_scaler)
    others = [(rid, ratio, compute_lh(o, case.full_schedule(rid, ratio), case.root_clv, case.root_scaler, mode=MODE_ENGINE))
              for rid, ratio in ((9, 0.5), (5, 0.8), (2 * n - 4, 0.1))]
    compute_lh(o, sched, case.root_clv, case.root_scaler)
    some = list(sched[0][:: max(1, len(sched[0]) // 24)]) + [sched[0][-1]]
    want_clv = {op.parent_clv_index: o.get_clv(op.parent_clv_index).copy() for op in some}
    want_sc = {op.parent_scaler_index: o.get_scaler(op.parent_scaler_index).copy() for op in some}
    for groups in (1, 0, 2, 3, 4, 8, 16):
        g = Partition(case.n, case.S, case.K)
        case.setup(g)
        g.set_subtree_groups(groups)
        g.reset_stats()
        got, ps = compute_lh(g, sched, case.root_clv, case.root_scaler, persite=True)
        assert same_bits([got], [want]) and same_bits(ps, want_ps), groups
        for op in some:
            assert same_bits(g.get_clv(op.parent_clv_index), want_clv[op.parent_clv_index]), groups
            assert np.array_equal(g.get_scaler(op.parent_scaler_index), want_sc[op.parent_scaler_index]), groups
        st = g.stats()
        assert st["program_launches"] == 1
        if groups or n >= 130:  # (the 64-taxon traversal is too short for the cost model to be sure)
            assert st["grouped_programs"] == (0 if groups == 1 else 1), (groups, st)
        by_default = st["grouped_programs"]
        # a streak: the second traversal onwards keeps most CLVs in registers, groups or not
        for it in range(4):
            assert same_bits([compute_lh(g, sched, case.root_clv, case.root_scaler)], [want]), (groups, it)
        st = g.stats()
        assert st["lazy_evaluations"] == 3 and st["grouped_programs"] == 5 * by_default, (groups, st)
        # ... and the state it leaves is the eager one (one replay on the first read)
        assert same_bits([compute_lh_root(g, ds, case.root_clv, case.root_scaler)], [want_root]), groups
        compute_lh(g, sched, case.root_clv, case.root_scaler)
        for op in some:
            assert same_bits(g.get_clv(op.parent_clv_index), want_clv[op.parent_clv_index]), groups
            assert np.array_equal(g.get_scaler(op.parent_scaler_index), want_sc[op.parent_scaler_index]), groups
        # the lowering of a traversal is kept and reused while the structure stays the same (new P slots
        # every time); another root is another program, and coming back finds the first one again
        if n >= 32:
            assert g.stats()["programs_reused"] >= 3, g.stats()
        for rid, ratio, w in others:
            assert same_bits([compute_lh(g, case.full_schedule(rid, ratio), case.root_clv, case.root_scaler)], [w]), (groups, rid)
        assert same_bits([compute_lh(g, sched, case.root_clv, case.root_scaler)], [want]), groups
        del g


def test_launch_configs_do_not_change_results():
    case = Case(25, 5000, 4, seed=11, data="ambiguous", weights="random")
    g, o = make(case)
    sched = case.full_schedule(3, 0.4)
    want = compute_lh(o, sched, case.root_clv, case.root_scaler, mode=MODE_ENGINE)
    for ctas, threads, elems in [(1, 64, 1), (2, 256, 2), (1, 384, 4), (2, 128, 1), (2, 96, 2), (1, 160, 4), (1, 512, 1)]:
        g.set_launch_config(ctas, threads, elems)
        for tail in (1, 2, 0):
            g.set_tail_mode(tail)
            got = compute_lh(g, sched, case.root_clv, case.root_scaler)
            assert same_bits([got], [want]), (ctas, threads, elems, tail)
            assert same_bits(g.get_clv(case.root_clv), o.get_clv(case.root_clv)), (ctas, threads, elems, tail)


def test_error_paths():
    from root_digger_b200.capi import EngineError, Partition
    with pytest.raises(EngineError):
        Partition(4, 10, 33)  # more rate categories than the engine carries
    p = Partition(4, 10, 4)
    with pytest.raises(EngineError):
        p.set_tip_states(0, b"ACGTACGT!J")  # illegal state code
    with pytest.raises(EngineError):
        p.update_prob_matrices([99], [0.1])
    with pytest.raises(EngineError):
        p.update_prob_matrices([0], [-1.0])
    # empty partition
    e = Partition(4, 0, 4)
    assert e.root_loglikelihood(6, 2) == 0.0
