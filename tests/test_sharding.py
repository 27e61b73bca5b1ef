"""N > 1 host logic on CPU: shard planners and a world_size-2 gloo run in which
every rank evaluates its site shard with the oracle and the per-shard tree nodes
are all-reduced exactly as the CUDA engine does over NCCL (zeros elsewhere, so
the sum is exact and independent of the number of shards)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from root_digger_b200 import sharding

ROOT = Path(__file__).resolve().parent.parent


def test_site_shard_plan():
    A = sharding.ALIGN
    assert A == 256  # include/rdk.h RDK_SHARD_ALIGN
    for S in (0, 1, 255, 256, 257, 1023, 1024, 1025, 100_000, 1_000_000):
        for G in (1, 2, 3, 8):
            plan = sharding.plan_site_shards(S, G)
            assert len(plan) == G and sum(c for _, c in plan) == S
            pos = 0
            for off, cnt in plan:
                assert off == pos and (off % A == 0 or cnt == 0)
                pos += cnt
    counts = [c for _, c in sharding.plan_site_shards(1_000_000, 8)]
    assert max(counts) - min(counts) <= 2 * A   # balanced to one alignment block (last one ragged)
    # cfg2 over 8 GPUs: 12 544 sites = 1 568 warp iterations per GPU fit one E = 4 pass of a sweep in
    # 4 chunks (37 CTAs x 11 warps x 4); the 13 312 of 1024-site blocks did not
    assert max(c for _, c in sharding.plan_site_shards(100_000, 8)) == 12_544


def test_grid_plan_prefers_replicas_that_fit():
    """cfg2 / cfg3 fit a full-site replica per GPU (roots distribute), cfg5 does not"""
    assert sharding.plan_grid(8, 500, 100_000) == (1, 8)
    assert sharding.plan_grid(8, 2000, 500_000) == (1, 8)
    assert sharding.plan_grid(8, 10_000, 1_000_000) == (8, 1)
    assert sharding.plan_grid(8, 5_000, 500_000) == (4, 2)
    assert sharding.plan_grid(8, 500, 100_000, force="sites") == (8, 1)
    assert sharding.plan_grid(1, 10_000, 1_000_000) == (1, 1)
    assert abs(sharding.replica_bytes(10_000, 125_000) - 171.2e9) < 0.1e9
    for n in (1, 2, 4, 8):
        gs, gr = sharding.plan_grid(n, 2000, 500_000, budget_bytes=40e9)
        assert gs * gr == n


def test_root_shard_plan_matches_reference_rule():
    """src/model.cpp:1899-1907: chunk*rank + min(mod, rank)"""
    assert sharding.plan_root_shards(range(17), 4) == [[0, 1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12], [13, 14, 15, 16]]
    assert sharding.plan_root_shards(range(3), 8)[3:] == [[]] * 5
    assert sum(sharding.plan_root_shards(range(19997), 8), []) == list(range(19997))
    assert sharding.plan_partition_shards(8, 8) == [[i] for i in range(8)]


def _worker(rank, world, port, S, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch
    import torch.distributed as dist
    from cases import Case, compute_lh
    from oracle_capi import MODE_ENGINE, OraclePartition, load_oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = Case(12, S, 4, seed=77, data="ambiguous", weights="random")
    off, cnt = sharding.plan_site_shards(S, world)[rank]
    sl = slice(off, off + cnt)
    part = OraclePartition(case.n, cnt, 4)
    case.setup(part, site_slice=sl)
    sched = case.full_schedule(2, 0.4)
    _, persite = compute_lh(part, sched, case.root_clv, case.root_scaler, persite=True, mode=MODE_ENGINE)
    # per-shard nodes of the canonical tree (one per block of RDK_SHARD_ALIGN sites), zeros elsewhere
    L = load_oracle()
    import ctypes as C
    A = sharding.ALIGN
    nblocks = (S + A - 1) // A
    nodes = np.zeros(nblocks)
    for b in range((cnt + A - 1) // A):
        seg = np.ascontiguousarray(persite[b * A:(b + 1) * A])
        pad = np.zeros(A)
        pad[:len(seg)] = seg
        nodes[off // A + b] = L.rdo_pairwise_sum(pad.ctypes.data_as(C.POINTER(C.c_double)), A)
    t = torch.from_numpy(nodes)
    dist.all_reduce(t)
    total = L.rdo_pairwise_sum(t.numpy().ctypes.data_as(C.POINTER(C.c_double)), nblocks)
    if rank == 0:
        q.put(total)
    dist.destroy_process_group()


@pytest.mark.parametrize("S", [5000, 2048])
def test_two_rank_gloo_site_sharding_matches_single_shard(S):
    import torch.multiprocessing as mp
    from cases import Case, compute_lh
    from oracle_capi import MODE_ENGINE, OraclePartition
    case = Case(12, S, 4, seed=77, data="ambiguous", weights="random")
    full = OraclePartition(case.n, S, 4)
    case.setup(full)
    want = compute_lh(full, case.full_schedule(2, 0.4), case.root_clv, case.root_scaler, mode=MODE_ENGINE)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, S, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.hex() == want.hex()   # bit-identical whatever the number of shards


def _partition_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    import fixtures
    import oracle_capi
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = fixtures.load("10.fasta")
    mine = [PARTS[p] for p in sharding.plan_partition_shards(len(PARTS), world)[rank]]
    tree = capi.RootedTree(path=str(fx["tree_path"]), lib=lib)
    m = capi.Model(tree, fx["alignment"], rate_cats=4, compress=True, seed=3, partitions=mine)
    m.initialize_partitions(uniform_freqs=False)
    for j, p in enumerate(sharding.plan_partition_shards(len(PARTS), world)[rank]):
        m.set_params(rates=_rates_of(p), alpha=0.5 + p, part=j)
    out = []
    for in_model in (False, True):  # the all-gather issued from Python / from inside model_t
        sm = sharding.PartitionShardedModel(m, len(PARTS), rank, world, dist, in_model=in_model)
        out += [sm.compute_lh(3, 0.4), sm.compute_lh_root(3, 0.7)] + sm.sweep_root_lh().tolist()
        sm.close()
    if rank == 0:
        q.put(np.array(out))
    dist.destroy_process_group()


PARTS = [(0, 300), (300, 520), (520, 1000)]


def _rates_of(p):
    """independent substitution rates per (global) partition; the initial rates of model_t are drawn
    from the model's RNG in local partition order (quirk B-10), so the comparison sets them"""
    base = np.array([.34, .42, .24, .74, .16, .88, .75, .54, .20, .06, .08, .41])
    return np.roll(base, p) * (1.0 + 0.1 * p)


def test_two_rank_gloo_partition_sharding_matches_single_process():
    """BASELINE cfg4 in miniature: 3 partitions dealt out to 2 ranks (p % 2), independent parameters
    per partition; compute_lh / compute_lh_root / the placement sweep equal the single-process
    3-partition model bit for bit (terms all-gathered, added in partition order)"""
    import torch.multiprocessing as mp
    import fixtures
    import oracle_capi
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    fx = fixtures.load("10.fasta")
    tree = capi.RootedTree(path=str(fx["tree_path"]), lib=lib)
    m = capi.Model(tree, fx["alignment"], rate_cats=4, compress=True, seed=3, partitions=PARTS)
    m.initialize_partitions(uniform_freqs=False)
    for p in range(len(PARTS)):
        m.set_params(rates=_rates_of(p), alpha=0.5 + p, part=p)
    want = np.array([m.compute_lh(3, 0.4), m.compute_lh_root(3, 0.7)] + m.sweep_root_lh().tolist())
    terms = m.last_sweep_partition_lh()
    assert terms.shape == (3, m.root_count) and np.array_equal((terms[0] + terms[1]) + terms[2], want[2:])
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_partition_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.concatenate([want, want])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


SEARCH_ARGS = dict(min_roots=1, root_ratio=0.05, atol=1e-3, pgtol=1e-3, brtol=1e-4, factor=1e12)
EXHAUSTIVE_ARGS = (1e-2, 1e-2, 1e-3, 1e13)


def _checkpoint_records(prefix, lib):
    """every record of "<prefix>.ckp" with the parameters of all of its partitions, as one flat array"""
    from root_digger_b200 import capi
    ck = capi.Checkpoint(prefix, lib=lib)
    rows = []
    for i, (rid, llh, alpha, nparts) in enumerate(ck.read_results()):
        rows += [float(rid), llh, alpha, float(nparts)]
        for part in range(nparts):
            pr = ck.read_params(i, part, K=2)
            rows += pr["rates"].tolist() + pr["freqs"].tolist() + [pr["alpha"]]
    ck.close()
    return np.array(rows)


def _whole_model_results(m, strategy):
    """what a run does with a multi-partition model, in order: initialisation (draws from the model's
    generator), slopes and a position search on a branch, a search from ranked / shuffled starts,
    exhaustive mode on a few branches"""
    out = []
    m.compute_lh(0, 0.5)
    for rid, x in ((2, 0.3), (5, 1.0), (7, 0.0)):
        m.compute_lh(rid, 0.5)
        out += list(m.compute_dlh(rid, x))
        out.append(m.optimize_alpha(rid, 0.5, 1e-9))
    rid, alpha, lh = m.search(strategy=strategy, **SEARCH_ARGS)
    out += [float(rid), alpha, lh]
    out += [m.compute_lh(rid, alpha)]  # with the parameters the best record put back
    m.set_max_outer_iterations(2)
    ids, llh, al = m.exhaustive_search(*EXHAUSTIVE_ARGS, rank=0, num_tasks=9)
    out += ids.astype(np.float64).tolist() + llh.tolist() + al.tolist() + m.lwr(llh).tolist()
    return np.array(out)


def _partition_search_worker(rank, world, port, strategy, q, ckp_prefix=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      OMP_NUM_THREADS="2")  # two ranks share the test machine
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    import fixtures
    import oracle_capi
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = fixtures.load("10.fasta")
    mine = [PARTS[p] for p in sharding.plan_partition_shards(len(PARTS), world)[rank]]
    tree = capi.RootedTree(path=str(fx["tree_path"]), lib=lib)
    m = capi.Model(tree, fx["alignment"], rate_cats=2, compress=True, seed=11, early_stop=True, partitions=mine)
    sm = sharding.PartitionShardedModel(m, len(PARTS), rank, world, dist)
    sm.initialize_partitions(uniform_freqs=False)
    if ckp_prefix:
        sm.set_checkpoint(ckp_prefix)  # one file per rank: "<prefix>.part<rank>of<world>.ckp"
    got = _whole_model_results(sm, strategy)
    q.put((rank, got, sm.exchanges))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("strategy", ["random", "modified_mad"])
def test_two_rank_gloo_partition_sharded_search_matches_single_process(strategy, tmp_path):
    """SURVEY 8e-3 for the WHOLE of model_t: with the sums over partitions completed inside model_t
    (set_partition_exchange), compute_dlh, optimize_alpha, search (ranked and shuffled starts -- the
    shuffle needs the ranks' generators in step with a single process's) and exhaustive mode + LWR on
    3 partitions dealt to 2 ranks return the bits of one process holding all three; every rank
    reports the same values"""
    import torch.multiprocessing as mp
    import fixtures
    import oracle_capi
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    fx = fixtures.load("10.fasta")
    tree = capi.RootedTree(path=str(fx["tree_path"]), lib=lib)
    m = capi.Model(tree, fx["alignment"], rate_cats=2, compress=True, seed=11, early_stop=True, partitions=PARTS)
    m.initialize_partitions(uniform_freqs=False)
    m.set_checkpoint(str(tmp_path / "one"))
    want = _whole_model_results(m, strategy)
    m.close()
    want_log = _checkpoint_records(str(tmp_path / "one"), lib)
    assert len(want_log) > 4 + 3 * 17  # at least one record carrying the parameters of the three partitions
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_partition_search_worker, args=(r, 2, port, strategy, q, str(tmp_path / "two")))
             for r in range(2)]
    for p in procs:
        p.start()
    import queue
    results, waited = {}, 0
    while len(results) < 2:
        try:
            r, v, n = q.get(timeout=5)
            results[r] = (v, n)
        except queue.Empty:
            waited += 5
            assert waited < 900 and all(p.exitcode in (None, 0) for p in procs), "a rank died"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        got, exchanges = results[r]
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (r, got, want)
        assert exchanges > 20
    assert results[0][1] == results[1][1], "both ranks took part in the same collectives"
    # every rank's log holds COMPLETE records -- the parameters of all three partitions, gathered
    # through the same exchange -- equal to the single process's log, byte for byte in content
    for r in (0, 1):
        got_log = _checkpoint_records(str(tmp_path / ("two.part%dof2" % r)), lib)
        assert np.array_equal(got_log.view(np.uint64), want_log.view(np.uint64)), r


def test_partition_exchange_argument_checks():
    """model_t::set_partition_exchange refuses layouts it cannot keep in step with a single process"""
    import ctypes as C
    import fixtures
    import oracle_capi
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    fx = fixtures.load("10.fasta")
    m = capi.Model(capi.RootedTree(path=str(fx["tree_path"]), lib=lib), fx["alignment"], rate_cats=2, partitions=PARTS[:2])

    class _Dist:  # never reached: the constructor must fail first
        pass

    with pytest.raises(ValueError, match="exactly this rank"):
        sharding.PartitionShardedModel(m, 5, 0, 2, _Dist())   # rank 0 of 2 holds 3 of 5 partitions, not 2
    with pytest.raises(ValueError, match="at most as many ranks"):
        sharding.PartitionShardedModel(m, 2, 0, 3, _Dist())
    fn_t = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_size_t, C.c_size_t, C.POINTER(C.c_double), C.c_void_p)
    m.L.rdh_model_set_partition_exchange.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.c_uint, C.c_uint, fn_t, C.c_void_p]
    cb = fn_t(lambda *a: None)
    for idx, total in (((1, 0), 4), ((0, 4), 4), ((0,), 4)):  # decreasing; out of range; wrong count
        arr = (C.c_uint * len(idx))(*idx)
        assert m.L.rdh_model_set_partition_exchange(m.h, arr, len(idx), total, cb, None) == 0
        assert b"global ind" in m.L.rdh_last_error() or b"one global index" in m.L.rdh_last_error()
    m.close()


def test_sweep_chunk_count_is_agreed_across_site_shards():
    """site-sharded runs add the ranks' sweep values slot by slot: every rank must cut the directed
    sweep into the same chunks.  The count therefore comes from the largest shard of the layout, not
    from the local site count -- here the two ranks hold 4096 and 3904 patterns of 8000, and the
    (test) hint would say 1 chunk for the first and 3 for the second"""
    import oracle_capi
    from cases import Case
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    S = 8000
    case = Case(12, S, 4, seed=3, data="iid")
    shards = sharding.plan_site_shards(S, 2)
    assert [c for _, c in shards] == [4096, 3904]
    got = []
    for off, cnt in shards:
        aln = {l: s[off:off + cnt] for l, s in case.aln.items()}
        m = capi.Model(capi.RootedTree(case.newick, lib=lib), aln, rate_cats=4, site_offset=off, global_sites=S,
                       nranks=2, rank=0 if off == 0 else 1)
        got.append(m.sweep_chunks)
        m.close()
    assert got == [1, 1]
    # unsharded models of those sizes do differ: the rule is what makes the ranks agree
    solo = [capi.Model(capi.RootedTree(case.newick, lib=lib), {l: s[:n] for l, s in case.aln.items()},
                       rate_cats=4).sweep_chunks for n in (4096, 3904)]
    assert solo == [1, 3]


def test_generator_state_round_trip_and_failed_initialisation_keeps_ranks_in_step():
    """what PartitionShardedModel.initialize_partitions relies on when a partition has no empirical
    frequencies: the model's generator can be read, put back and advanced (std::minstd_rand is one
    integer), and an initialisation that throws half way leaves it as many draws further as the
    partitions it got through -- so the ranks that did NOT fail can be put where the failing one is"""
    import ctypes as C
    import fixtures
    import oracle_capi
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    fx = fixtures.load("10.fasta")

    def model(aln, parts):
        m = capi.Model(capi.RootedTree(path=str(fx["tree_path"]), lib=lib), aln, rate_cats=1, seed=99, partitions=parts)
        m.L.rdh_model_rng_state.argtypes = [C.c_void_p]
        m.L.rdh_model_rng_state.restype = C.c_ulonglong
        m.L.rdh_model_set_rng_state.argtypes = [C.c_void_p, C.c_ulonglong]
        m.L.rdh_model_set_rng_state.restype = None
        m.L.rdh_model_discard_rng.argtypes = [C.c_void_p, C.c_ulonglong]
        m.L.rdh_model_discard_rng.restype = None
        m.L.rdh_model_first_partition_without_empirical_freqs.argtypes = [C.c_void_p]
        return m

    m = model(fx["alignment"], PARTS)
    L, h = m.L, m.h
    s0 = L.rdh_model_rng_state(h)
    first = m.assign_indicies("search", 5, 0.0, 0, 1, strategy="random")
    assert L.rdh_model_rng_state(h) != s0
    L.rdh_model_set_rng_state(h, s0)
    assert m.assign_indicies("search", 5, 0.0, 0, 1, strategy="random") == first        # same state, same shuffle
    L.rdh_model_set_rng_state(h, s0)
    L.rdh_model_discard_rng(h, 1)
    assert m.assign_indicies("search", 5, 0.0, 0, 1, strategy="random") != first
    # one draw per partition: initialize_partitions advances the generator by exactly len(PARTS) draws
    L.rdh_model_set_rng_state(h, s0)
    m.initialize_partitions(uniform_freqs=False)
    after_init = L.rdh_model_rng_state(h)
    L.rdh_model_set_rng_state(h, s0)
    L.rdh_model_discard_rng(h, len(PARTS))
    assert L.rdh_model_rng_state(h) == after_init
    assert L.rdh_model_first_partition_without_empirical_freqs(h) == -1
    m.close()

    # no G in the columns of the middle partition: its empirical frequency of G is zero
    aln = {}
    for label, seq in fx["alignment"].items():
        seq = seq if isinstance(seq, str) else seq.decode()
        b, e = PARTS[1]
        aln[label] = seq[:b] + seq[b:e].replace("G", "A").replace("g", "a") + seq[e:]
    bad = model(aln, PARTS)
    L, h = bad.L, bad.h
    assert L.rdh_model_first_partition_without_empirical_freqs(h) == 1
    s0 = L.rdh_model_rng_state(h)
    with pytest.raises(RuntimeError, match="frequen"):
        bad.initialize_partitions(uniform_freqs=False)
    failed_at = L.rdh_model_rng_state(h)
    L.rdh_model_set_rng_state(h, s0)
    L.rdh_model_discard_rng(h, 1)            # partition 0 drew, partition 1 threw before its draw
    assert L.rdh_model_rng_state(h) == failed_at
    bad.initialize_partitions(uniform_freqs=True)   # what RootDigger's main falls back to
    bad.close()


def _no_g_in_partition(aln, part):
    out = {}
    for label, seq in aln.items():
        seq = seq if isinstance(seq, str) else seq.decode()
        b, e = part
        out[label] = seq[:b] + seq[b:e].replace("G", "A").replace("g", "a") + seq[e:]
    return out


def _starts_after_fallback(m):
    """RootDigger's main (src/main.cpp:560-589): empirical frequencies, uniform ones if they are invalid;
    then the shuffled starts of a search and one evaluation"""
    try:
        m.initialize_partitions(uniform_freqs=False)
        fell_back = 0.0
    except RuntimeError:
        m.initialize_partitions(uniform_freqs=True)
        fell_back = 1.0
    starts = m.assign_indicies("search", 6, 0.0, 0, 1, strategy="random")
    return np.array([fell_back] + [float(s) for s in starts] + [m.compute_lh(int(starts[0]), 0.5)])


def _fallback_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      OMP_NUM_THREADS="2")
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    import torch.distributed as dist
    import fixtures
    import oracle_capi
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = fixtures.load("10.fasta")
    aln = _no_g_in_partition(fx["alignment"], PARTS[1])
    mine = [PARTS[p] for p in sharding.plan_partition_shards(len(PARTS), world)[rank]]
    m = capi.Model(capi.RootedTree(path=str(fx["tree_path"]), lib=lib), aln, rate_cats=1, seed=99, partitions=mine)
    sm = sharding.PartitionShardedModel(m, len(PARTS), rank, world, dist)
    q.put((rank, _starts_after_fallback(sm)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_failed_empirical_frequencies_fail_on_every_rank_and_keep_them_in_step():
    """partition 1 (held by rank 1) has no G: a single process throws when its initialisation gets
    there, falls back to uniform frequencies and shuffles its starts with a generator that is one
    draw (partition 0's) + three draws further.  On shards BOTH ranks must raise -- rank 0 holds only
    healthy partitions -- and end up with the same starts and the same likelihood as that process"""
    import torch.multiprocessing as mp
    import fixtures
    import oracle_capi
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    fx = fixtures.load("10.fasta")
    aln = _no_g_in_partition(fx["alignment"], PARTS[1])
    m = capi.Model(capi.RootedTree(path=str(fx["tree_path"]), lib=lib), aln, rate_cats=1, seed=99, partitions=PARTS)
    want = _starts_after_fallback(m)
    m.close()
    assert want[0] == 1.0
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fallback_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert np.array_equal(results[r].view(np.uint64), want.view(np.uint64)), (r, results[r], want)


def test_partition_exchange_failures_surface_and_detach_restores_the_local_model():
    """a failing all-gather cannot raise through the C++ frames: the sums are poisoned, model_t refuses
    the NaN and the wrapper reports the cause; close() detaches the exchange; without the in-model
    exchange the calls that need global sums inside model_t are refused"""
    import fixtures
    import oracle_capi
    import oracle_build
    from root_digger_b200 import capi
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    lib = capi.load_tree_lib(oracle_build.build_host_on_oracle())
    fx = fixtures.load("10.fasta")
    m = capi.Model(capi.RootedTree(path=str(fx["tree_path"]), lib=lib), fx["alignment"], rate_cats=1, seed=3,
                   partitions=PARTS[:2])
    m.initialize_partitions(uniform_freqs=False)
    alone = m.compute_lh(2, 0.5)

    class Broken:
        def all_gather(self, out, buf):
            raise OSError("the network is down")

    class Loopback:  # one rank: the gathered terms are the local ones
        def all_gather(self, out, buf):
            out[0].copy_(buf)

    sm = sharding.PartitionShardedModel(m, 2, 0, 1, Broken())
    with pytest.raises(RuntimeError, match="partition exchange failed.*network is down"):
        sm.compute_lh(2, 0.5)
    m2 = capi.Model(capi.RootedTree(path=str(fx["tree_path"]), lib=lib), fx["alignment"], rate_cats=1, seed=3,
                    partitions=PARTS[:2])
    m2.initialize_partitions(uniform_freqs=False)
    ok = sharding.PartitionShardedModel(m2, 2, 0, 1, Loopback())
    assert ok.compute_lh(2, 0.5).hex() == alone.hex() and ok.exchanges == 1
    ok.set_params(1, rates=_rates_of(1))   # a GLOBAL partition index
    ok.set_params(5, rates=_rates_of(2))   # not held here: ignored
    changed = ok.compute_lh(2, 0.5)
    assert changed != alone and ok.exchanges == 2
    ok.close()
    assert m2.compute_lh(2, 0.5).hex() == changed.hex() and ok.exchanges == 2   # detached: no further exchange
    m2.close()
    sm.close()   # several wrappers alive at once, closed in any order
    assert m.compute_lh(2, 0.5).hex() == alone.hex()
    m.close()

    m3 = capi.Model(capi.RootedTree(path=str(fx["tree_path"]), lib=lib), fx["alignment"], rate_cats=1, seed=3,
                    partitions=PARTS[:2])
    m3.initialize_partitions(uniform_freqs=False)
    outside = sharding.PartitionShardedModel(m3, 2, 0, 1, Loopback(), in_model=False)
    assert outside.compute_lh(2, 0.5).hex() == alone.hex()
    with pytest.raises(RuntimeError, match="in_model=True"):
        outside.optimize_alpha(2, 0.5, 1e-7)
    m3.close()
