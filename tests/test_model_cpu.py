"""model_t mirror (host C++) on the ORACLE backend: the reference's own model_t
tests (test/src/model.cpp) restated.  The same host sources compiled against the
CUDA engine are compared with this build in test_gpu_model.py."""
import math

import numpy as np
import pytest

import fixtures
import oracle_capi
import oracle_build
from root_digger_b200 import capi


@pytest.fixture(scope="module")
def lib():
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    return capi.load_tree_lib(oracle_build.build_host_on_oracle())


def make_model(lib, name="10.fasta", K=1, seed=12345, uniform=True, **kw):
    fx = fixtures.load(name)
    tree = capi.RootedTree(path=str(fx["tree_path"]), lib=lib)
    m = capi.Model(tree, fx["alignment"], rate_cats=K, compress=True, invariant_sites=True, seed=seed, **kw)
    m.initialize_partitions(uniform_freqs=uniform)
    return m


def test_pattern_compression_matches_reference_counts(lib):
    assert make_model(lib, "10.fasta").sites() == 991
    assert make_model(lib, "101.phy").sites() == 1630


def test_compute_lh_invariants(lib):
    """test/src/model.cpp:59-75, :271-288"""
    m = make_model(lib)
    assert m.root_count == 17
    for rid in range(m.root_count):
        a, b, c = m.compute_lh(rid), m.compute_lh(rid), m.compute_lh_root(rid)
        assert math.isfinite(a) and a < 0 and a == b and a == c


def test_compute_dlh_and_optimize_alpha(lib):
    """test/src/model.cpp:95-110, :132-238"""
    m = make_model(lib)
    for rid in range(m.root_count):
        m.compute_lh(rid)
        lh, dlh = m.compute_dlh(rid)
        assert math.isfinite(lh) and math.isfinite(dlh)
        for start in (0.5, 0.0, 1.0):
            m.compute_lh(rid, start)
            r = m.optimize_alpha(rid, start, 1e-7)
            assert 0.0 <= r <= 1.0


def test_batched_probes_have_the_bits_of_the_reference_call_sequence(lib):
    """compute_dlh (1 fused call for both evaluations) and optimize_alpha (5 evaluations
    before the first decision, one call per dyadic level, one per Brent iteration) against the same
    model issuing every root-only evaluation on its own, the way src/model.cpp:481-519,679-794 does
    -- every root, three start positions, loose and tight tolerances; one and several partitions"""
    from root_digger_b200 import synth
    top = synth.random_tree(12, 4)
    rates, freqs = synth.random_params(9)
    aln = synth.simulate_alignment(top, 900, 5, rates, freqs, capi.gamma_cats(1.0, 2))

    def pair(**kw):
        out = []
        for batched in (True, False):
            if kw:
                m = capi.Model(capi.RootedTree(synth.to_newick(top), lib=lib), aln, rate_cats=2, compress=True, seed=3,
                               **kw)
                m.initialize_partitions(uniform_freqs=False)
            else:
                m = make_model(lib, K=2, uniform=False)
            m.set_batched_probes(batched)
            assert m.batched_probes is batched
            out.append(m)
        return out

    for a, b in (pair(), pair(partitions=[(0, 300), (300, 900)])):
        for rid in range(a.root_count):
            for m in (a, b):
                m.compute_lh(rid)
            for x in (0.0, 0.37, 1.0 - 1e-9, 1.0):
                assert [v.hex() for v in a.compute_dlh(rid, x)] == [v.hex() for v in b.compute_dlh(rid, x)]
            for start, atol in ((0.5, 1e-7), (0.0, 1e-14), (1.0, 1e-3)):
                assert a.optimize_alpha(rid, start, atol).hex() == b.optimize_alpha(rid, start, atol).hex()
                # what follows an optimize_alpha in both drivers: the root-only evaluation there
                assert a.compute_lh_root(rid, 0.3).hex() == b.compute_lh_root(rid, 0.3).hex()
        # the batched model did batch (a slope is one call of 2, the first decision of optimize_alpha
        # one call of 5); the other issued every evaluation on its own
        ca, cb = a.probe_counters(), b.probe_counters()
        assert ca["fused_batches"] > 0 and ca["fused_evaluations"] >= 2 * ca["fused_batches"]
        # (evaluations made ahead of need can only add to the batched count)
        assert cb["fused_batches"] == 0 and cb["single_evaluations"] <= ca["single_evaluations"] + ca["fused_evaluations"]
        a.close()
        b.close()


def test_optimize_root_location(lib):
    """test/src/model.cpp:240-252"""
    m = make_model(lib)
    m.compute_lh(0)
    rid, alpha, lh = m.optimize_root_location(1, .05)
    assert 0.0 <= alpha <= 1.0 and math.isfinite(lh)


def test_fused_sweep_equals_reference_loop(lib):
    m = make_model(lib, K=4)
    m.compute_lh(0)
    m.set_fused(True)
    a = m.sweep_root_lh()
    m.compute_lh(0)
    m.set_fused(False)
    b = m.sweep_root_lh()
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64))


@pytest.mark.parametrize("name,K", [("10.fasta", 1), ("10.fasta", 4), ("101.phy", 4)])
def test_directed_sweep_has_the_bits_of_the_reference_loop(lib, name, K):
    """the three sweep modes -- the reference's move_root + compute_lh_root loop
    (src/model.cpp:871-874), the same operations as one engine call, and the directed-CLV
    pre-order pass -- give bit-identical log-likelihoods from ANY current root, and the
    directed pass leaves the model where it was (same root, same compute_lh_root)"""
    m = make_model(lib, name, K=K, uniform=(name != "101.phy"))
    nroots = m.root_count
    for start in sorted({0, 3, nroots // 2, nroots - 1}):
        out = []
        for mode in (m.SWEEP_SEQUENTIAL, m.SWEEP_PATH, m.SWEEP_DIRECTED):
            lh0 = m.compute_lh(start, 0.3)
            m.set_sweep_mode(mode)
            out.append(m.sweep_root_lh())
        assert np.array_equal(out[0].view(np.uint64), out[1].view(np.uint64)), start
        assert np.array_equal(out[0].view(np.uint64), out[2].view(np.uint64)), start
        assert m.compute_lh_root(start, 0.3) == lh0           # directed: state untouched
        # two directed sweeps in a row, and a directed chunk
        again = m.sweep_root_lh()
        assert np.array_equal(again.view(np.uint64), out[2].view(np.uint64))
        part = m.sweep_root_lh(2, 7)
        assert np.array_equal(part.view(np.uint64), out[2][2:7].view(np.uint64))


def test_root_sharded_sweep_equals_full_sweep(lib):
    """root placements distributed over ranks (src/model.cpp:1899-1907 rule): every rank
    starts from the CLVs of root 0 and sweeps its own chunk; the chunks concatenate to
    the single-rank sweep bit for bit"""
    from root_digger_b200.sharding import plan_root_shards
    m = make_model(lib, K=4)
    m.compute_lh(0)
    full = m.sweep_root_lh()
    for nranks in (2, 3, 8):
        parts = []
        for chunk in plan_root_shards(range(m.root_count), nranks):
            m.compute_lh(0)
            parts.append(m.sweep_root_lh(chunk[0], chunk[-1] + 1) if chunk else np.zeros(0))
        got = np.concatenate(parts)
        assert np.array_equal(got.view(np.uint64), full.view(np.uint64)), nranks
    assert len(m.sweep_root_lh(5, 5)) == 0
    with pytest.raises(Exception):
        m.sweep_root_lh(3, m.root_count + 1)


def test_move_root_invariance_under_jc(lib):
    """test/src/model.cpp:367-387"""
    m = make_model(lib, "101.phy", uniform=False)
    m.set_params(rates=np.ones(12), freqs=np.full(4, .25))
    m.compute_lh(0)
    v = m.sweep_root_lh()
    assert np.ptp(v) <= 1.19e-5 * abs(v[0])


def test_search_returns_consistent_likelihood(lib):
    """test/src/model.cpp:310-346: compute_lh(final_rl) == Approx(final_lh).

    Quirk kept from the reference ("B-17"): set_model_params (src/model.cpp:1913-1923)
    re-installs the optimiser's RAW frequency vector with set_freqs, while every likelihood during the search
    was computed with set_freqs_all_free (:350-355), i.e. the vector divided by its sum.  The recorded
    likelihood is therefore reproduced after normalising the installed frequencies."""
    m = make_model(lib)
    m.compute_lh(0)
    initial = m.compute_lh(4)
    rid, alpha, lh = m.search(3, 0.0, 1e-3, 1e-3, 1e-3, 1e12, strategy="random")
    assert lh >= initial
    rates, freqs, _ = m.get_params()
    m.set_params(freqs=freqs / freqs.sum())
    assert m.compute_lh(rid, alpha) == pytest.approx(lh, rel=1.2e-5)
    assert m.compute_lh_root(rid, alpha) == pytest.approx(lh, rel=1.2e-5)


def test_lwr_is_a_softmax(lib):
    m = make_model(lib)
    llh = np.array([-100.0, -101.0, -130.0, -100.5])
    w = m.lwr(llh)
    assert abs(w.sum() - 1) < 1e-15 and np.argmax(w) == 0
    assert np.allclose(w, np.exp(llh + 100) / np.exp(llh + 100).sum())


def test_taxa_mismatch_is_rejected(lib):
    fx = fixtures.load("10.fasta")
    tree = capi.RootedTree(path=str(fx["tree_path"]), lib=lib)
    aln = dict(fx["alignment"])
    aln["zzz"] = aln.pop("a")
    with pytest.raises(RuntimeError, match="inconsistient"):
        capi.Model(tree, aln, rate_cats=1)


@pytest.mark.slow
def test_exhaustive_search_runs(lib):
    """test/src/model.cpp:389-401"""
    m = make_model(lib, uniform=False)
    m.compute_lh(0)
    ids, llh, alpha = m.exhaustive_search(1e-3, 1e-3, 1e-3, 1e12)
    assert sorted(ids.tolist()) == list(range(17)) and np.isfinite(llh).all()
    assert ((alpha >= 0) & (alpha <= 1)).all()
    assert "LWR=" in m.newick()


def test_bench_exhaustive_sample_helper(lib):
    """bench.py --exhaustive-branches: the bounded exhaustive-mode sample on the oracle
    backend -- the requested number of branches is optimised, nothing else"""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("bench", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    m = make_model(lib, uniform=False)
    m.compute_lh(0)
    out = bench.exhaustive_sample(m, 2, tol=(1e-2, 1e-2, 1e-2, 1e13))
    assert 1 <= out["branches"] <= 2 and out["branches_per_sec"] > 0 and math.isfinite(out["best_llh"])
    assert out["best_branch"] in (0, 1)
    # the alpha loop is reported with batched and one-by-one root evaluations, which must agree
    assert out["optimize_alpha_same_bits"] is True and m.batched_probes
    assert out["us_per_compute_dlh"] > 0 and out["us_per_compute_dlh_unbatched"] > 0
