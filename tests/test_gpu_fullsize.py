"""GPU property tests at BASELINE.json's full single-GPU size (configs[1]: 500 taxa x
100 000 sites x 4 categories) and at the per-GPU shard shape of configs[2]
(2 000 taxa x 62 500 sites).  The oracle needs minutes at these sizes, so the
checks here are the size-independent properties of the domain:

  * the reference's invariants (test/src/model.cpp:59-75, 271-288, 367-387):
    finite, negative, bit-reproducible; full-traversal logL == root-only logL;
    all roots equal under a reversible model;
  * a placement scored by the sweep (move_root along a path + compute_lh_root)
    == the same placement scored by a full traversal from scratch, bit for bit
    (a CLV is a function of its subtree only, whatever order it was built in);
  * returning to the first root after the whole sweep reproduces its logL;
  * the total is the canonical pairwise tree over the per-site values (a
    checksum of checksums), and two half-alignment partitions produce exactly
    the two halves of the per-site vector (site shards are independent);
  * a sampled window of sites agrees with the CPU oracle bit for bit
    (columns are independent, so a 512-site slice of the big case is a small case).
"""
import numpy as np
import pytest

from cases import Case, compute_lh, compute_lh_root, move_root, same_bits
from oracle_capi import MODE_ENGINE, OraclePartition

pytestmark = pytest.mark.gpu


def canonical_tree_sum(x):
    """balanced binary tree over the zero-padded power-of-two index space"""
    n = 1
    while n < len(x):
        n *= 2
    a = np.zeros(n)
    a[:len(x)] = x
    while len(a) > 1:
        a = a[0::2] + a[1::2]
    return float(a[0])


@pytest.fixture(scope="module")
def cfg2():
    from root_digger_b200.capi import Partition, gamma_cats
    case = Case(500, 100000, 4, seed=0x5EED0002, data="iid", alpha=1.0, gamma_cats=gamma_cats)
    lay = case.tree.sweep_layout()
    g = Partition(case.n, case.S, case.K, clv_buffers=lay["clv_buffers"], scale_buffers=lay["scale_buffers"],
                  prob_matrices=lay["prob_matrices"])
    case.setup(g)
    yield case, g
    g.close()


def test_cfg2_full_evaluation_invariants(cfg2):
    case, g = cfg2
    sched = case.full_schedule(0, 0.5)
    lh, persite = compute_lh(g, sched, case.root_clv, case.root_scaler, persite=True)
    assert np.isfinite(lh) and lh < 0
    assert np.all(np.isfinite(persite)) and np.all(persite < 0)
    # bit-reproducible (test/src/model.cpp:59-75)
    assert same_bits([lh], [compute_lh(g, sched, case.root_clv, case.root_scaler)])
    # full traversal == root-only evaluation (test/src/model.cpp:271-288, zero tolerance)
    assert same_bits([lh], [compute_lh_root(g, case.derivative_schedule(0, 0.5), case.root_clv, case.root_scaler)])
    # checksum of checksums: the total is the canonical tree over the per-site values
    assert same_bits([lh], [canonical_tree_sum(persite)])
    # 500 iid taxa underflow 2^-256 many times over: the scalers must have fired
    assert g.get_scaler(case.root_scaler).min() >= 1


def test_cfg2_sampled_sites_match_oracle(cfg2):
    case, g = cfg2
    sched = case.full_schedule(0, 0.5)
    _, persite = compute_lh(g, sched, case.root_clv, case.root_scaler, persite=True)
    for lo in (0, 49920, 100000 - 512):
        sl = slice(lo, lo + 512)
        o = OraclePartition(case.n, 512, case.K)
        case.setup(o, sl)
        compute_lh(o, sched, case.root_clv, case.root_scaler)
        _, want = o.root_loglikelihood(case.root_clv, case.root_scaler, persite=True, mode=MODE_ENGINE)
        assert same_bits(persite[sl], want), lo
        assert np.array_equal(g.get_scaler(case.root_scaler)[sl], o.get_scaler(case.root_scaler))
        o.close()


def test_cfg2_sweep_is_path_independent(cfg2):
    case, g = cfg2
    nroots = case.tree.root_count
    assert nroots == 2 * case.n - 3
    lh0 = compute_lh(g, case.full_schedule(0, 0.5), case.root_clv, case.root_scaler)
    roots = list(range(nroots)) + [0]
    sw = case.sweep_schedule(roots, 0.5)
    out = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler)
    assert np.all(np.isfinite(out)) and np.all(out < 0)
    # back at the first root after visiting every branch
    assert same_bits([out[0]], [lh0]) and same_bits([out[-1]], [lh0])
    # placements scored along the sweep == scored by a full traversal from scratch
    for rid in (1, nroots // 3, nroots - 1):
        fresh = compute_lh(g, case.full_schedule(rid, 0.5), case.root_clv, case.root_scaler)
        assert same_bits([out[rid]], [fresh]), rid
    # a second sweep in the opposite order gives the same values
    case.tree.root_by(0, 0.5)
    compute_lh(g, case.full_schedule(0, 0.5), case.root_clv, case.root_scaler)
    rev = list(range(nroots - 1, -1, -1))
    out_rev = g.sweep_root_placements(*case.sweep_schedule(rev, 0.5), case.root_clv, case.root_scaler)
    assert same_bits(out_rev[::-1], out[:nroots])


def test_cfg2_directed_sweep_equals_path_sweep(cfg2):
    """at full size: the directed-CLV pre-order pass == the reference-shaped sweep, bit for bit, from
    two different current roots; its chunks concatenate to the whole; nothing it does is visible in
    the partition afterwards"""
    from root_digger_b200.capi import RDK_SWEEP_KEEP_ROOT
    case, g = cfg2
    lay = case.tree.sweep_layout()
    nroots = case.tree.root_count
    compute_lh(g, case.full_schedule(0, 0.5), case.root_clv, case.root_scaler)
    want = g.sweep_root_placements(*case.sweep_schedule(list(range(nroots)), 0.5), case.root_clv, case.root_scaler)
    for start in (0, nroots // 2):
        lh0 = compute_lh(g, case.full_schedule(start, 0.5), case.root_clv, case.root_scaler)
        *sw, pos = case.tree.generate_sweep_operations(layout=lay)
        assert len(pos) == nroots and len(sw[4]) == 2 * nroots - 1
        got = np.empty(nroots)
        got[pos] = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler, flags=RDK_SWEEP_KEEP_ROOT)
        assert same_bits(got, want), start
        assert same_bits([lh0], [compute_lh_root(g, case.derivative_schedule(start, 0.5), case.root_clv,
                                                 case.root_scaler)])
    # root-placement chunks (exhaustive mode's distribution over ranks)
    compute_lh(g, case.full_schedule(0, 0.5), case.root_clv, case.root_scaler)
    parts = []
    for lo, hi in ((0, 250), (250, 251), (251, nroots)):
        *sw, pos = case.tree.generate_sweep_operations(lo, hi, layout=lay)
        out = np.empty(hi - lo)
        out[pos - lo] = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler, flags=RDK_SWEEP_KEEP_ROOT)
        parts.append(out)
    assert same_bits(np.concatenate(parts), want)


def test_cfg2_site_shards_are_independent(cfg2):
    """two half-alignment partitions == the two halves of the per-site vector"""
    from root_digger_b200.capi import Partition
    from root_digger_b200.sharding import plan_site_shards
    case, g = cfg2
    sched = case.full_schedule(7, 0.3)
    _, persite = compute_lh(g, sched, case.root_clv, case.root_scaler, persite=True)
    for off, cnt in plan_site_shards(case.S, 2):
        h = Partition(case.n, cnt, case.K)
        case.setup(h, slice(off, off + cnt))
        _, ps = compute_lh(h, sched, case.root_clv, case.root_scaler, persite=True)
        assert same_bits(ps, persite[off:off + cnt])
        h.close()


def test_cfg2_root_invariance_under_reversible_model():
    """test/src/model.cpp:367-387 at full size: all-ones rates + uniform pi"""
    from root_digger_b200.capi import Partition, gamma_cats
    case = Case(500, 100000, 4, seed=77, data="iid", alpha=1.0, gamma_cats=gamma_cats)
    case.rates = np.ones(12)
    case.freqs = np.full(4, 0.25)
    g = Partition(case.n, case.S, case.K)
    case.setup(g)
    compute_lh(g, case.full_schedule(0, 0.5), case.root_clv, case.root_scaler)
    roots = list(range(case.tree.root_count))
    out = g.sweep_root_placements(*case.sweep_schedule(roots, 0.5), case.root_clv, case.root_scaler)
    assert np.ptp(out) <= 1e-9 * abs(out[0])
    g.close()


def test_cfg3_shard_shape():
    """configs[2] per-GPU shard: 2 000 taxa x 62 500 sites (16 GB of CLVs)"""
    from root_digger_b200.capi import Partition, gamma_cats
    case = Case(2000, 62500, 4, seed=0x5EED0003, data="iid", alpha=1.0, gamma_cats=gamma_cats)
    g = Partition(case.n, case.S, case.K)
    case.setup(g)
    sched = case.full_schedule(0, 0.5)
    lh, persite = compute_lh(g, sched, case.root_clv, case.root_scaler, persite=True)
    assert np.isfinite(lh) and lh < 0
    assert same_bits([lh], [canonical_tree_sum(persite)])
    assert same_bits([lh], [compute_lh_root(g, case.derivative_schedule(0, 0.5), case.root_clv, case.root_scaler)])
    # move the root to the far end of the id range and come back
    far = case.tree.root_count - 1
    move_root(g, case.move_schedule(far, 0.5))
    a = compute_lh_root(g, case.derivative_schedule(far, 0.5), case.root_clv, case.root_scaler)
    b = compute_lh(g, case.full_schedule(far, 0.5), case.root_clv, case.root_scaler)
    assert same_bits([a], [b])
    move_root(g, case.move_schedule(0, 0.5))
    assert same_bits([lh], [compute_lh_root(g, case.derivative_schedule(0, 0.5), case.root_clv, case.root_scaler)])
    o = OraclePartition(case.n, 256, case.K)
    case.setup(o, slice(31000, 31256))
    compute_lh(o, sched, case.root_clv, case.root_scaler)
    _, want = o.root_loglikelihood(case.root_clv, case.root_scaler, persite=True, mode=MODE_ENGINE)
    assert same_bits(persite[31000:31256], want)
    g.close()
