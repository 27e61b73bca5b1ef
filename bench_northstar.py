"""North-star configurations of BASELINE.json beyond the headline (bench.py imports this).

  cfg3  synthetic 2 000-taxon x 500 000-site DNA, exhaustive LWR over all branches, site-sharded
        across the GPUs (62 500 sites per GPU at N = 8)
  cfg4  8 partitions x 50 000 sites with independent non-reversible models, one partition per GPU,
        scalar sum over partitions (reference src/model.cpp:397,429)
  cfg5  synthetic 10 000-taxon x 1 000 000-site alignment, exhaustive mode, site-sharded
        (125 000 sites and ~165 GB of CLVs per GPU at N = 8)

Every configuration runs through the reference-facing host API (librd_host.so `model_t`:
compute_lh, the placement sweep of suggest_roots_lh, exhaustive_search) with host buffers, one
process per GPU, site shards joined by the engine's own NCCL all-reduce.  What is reported per
configuration:

  full_evaluation   model_t::compute_lh: wall ms (max over ranks), device ms of the program kernel,
                    algorithmic GB/s per GPU and its fraction of the measured HBM peak
  sweep             all 2n-3 placements (suggest_roots_lh's loop as one directed pass): placements/s
  oracle_window     per-site log-likelihoods of a window of this rank's sites against the CPU oracle
                    on the same columns: bit-identical in the engine's arithmetic and <= 1e-9 relative
                    against the reference-order arithmetic (the oracle is the checker, nothing timed)
  digest            SHA-256 of the bit patterns of all placement log-likelihoods (equal on every rank)
  exhaustive        (cfg3) a bounded sample of exhaustive mode: branches fully optimised per second,
                    full evaluations per second inside BFGS, microseconds per compute_dlh

The alignments are generated per rank for its own shard only (no rank ever holds the 10 GB of
cfg5): iid-uniform columns block by block from a seed that depends on the global block index, or
columns evolved down the tree for the shard.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import time

import numpy as np

CONFIGS = {
    # name: taxa, sites, rate categories, data, partitions
    "cfg3": dict(taxa=2000, sites=500_000, cats=4, data="evolved", partitions=1,
                 workload="cfg3: synthetic 2000-taxon x 500000-site DNA, UNREST+G4, exhaustive LWR over all "
                          "branches, site-sharded"),
    "cfg4": dict(taxa=500, sites=400_000, cats=4, data="evolved", partitions=8,
                 workload="cfg4: multi-partition MSA, 8 partitions x 50000 sites, independent UNREST+G4 models, "
                          "one partition per GPU"),
    "cfg5": dict(taxa=10_000, sites=1_000_000, cats=4, data="iid", partitions=1,
                 workload="cfg5: synthetic 10000-taxon x 1000000-site DNA, UNREST+G4, exhaustive mode, "
                          "site-sharded"),
}
# the headline workload at constant work per GPU (100 000 sites of 500 taxa on every GPU): what the
# strong-scaling curve of the headline cannot show -- it leaves 12.5 k sites per GPU at N = 8
CONFIGS["cfg2_weak"] = dict(taxa=500, sites=100_000, sites_per_gpu=True, cats=4, data="evolved", partitions=1,
                            workload="cfg2 at constant work per GPU: 500 taxa x (100000 x N) sites, UNREST+G4, "
                                     "site-sharded (weak scaling of the headline step)")
BLOCK = 1024  # site block of the data generator (a multiple of RDK_SHARD_ALIGN: a shard is a range of its sites)


def scaled(cfg: dict, scale: float) -> dict:
    if scale >= 1.0:
        return dict(cfg)
    out = dict(cfg)
    out["taxa"] = max(16, int(cfg["taxa"] * scale))
    out["sites"] = max(8 * BLOCK * cfg["partitions"], int(cfg["sites"] * scale) // (BLOCK * cfg["partitions"])
                       * BLOCK * cfg["partitions"])
    out["workload"] = cfg["workload"] + " [scaled x%g: %d taxa x %d sites]" % (scale, out["taxa"], out["sites"])
    return out


def iid_shard(n: int, off: int, cnt: int, seed: int) -> np.ndarray:
    """[n, cnt] uint8 ACGT, iid-uniform, a function of the GLOBAL site index only"""
    dna = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = np.empty((n, cnt), dtype=np.uint8)
    b0, b1 = off // BLOCK, (off + cnt + BLOCK - 1) // BLOCK
    for b in range(b0, b1):
        rng = np.random.default_rng([seed, b])
        blk = dna[rng.integers(0, 4, (n, BLOCK), dtype=np.uint8)]
        lo, hi = max(off, b * BLOCK), min(off + cnt, (b + 1) * BLOCK)
        out[:, lo - off:hi - off] = blk[:, lo - b * BLOCK:hi - b * BLOCK]
    return out


def evolved_shard(top, labels, off: int, cnt: int, seed: int, rates, freqs, cat_rates) -> np.ndarray:
    from root_digger_b200 import synth
    aln = synth.simulate_alignment(top, cnt, seed + 7919 * (off // BLOCK + 1), rates, freqs, cat_rates)
    return np.stack([np.frombuffer(aln[l], dtype=np.uint8) for l in labels])


class Setup:
    """tree, model parameters and this rank's columns of one configuration"""

    def __init__(self, cfg: dict, seed: int, off: int, cnt: int, part_seed: int = 0):
        from root_digger_b200 import capi, synth
        self.n, self.K = cfg["taxa"], cfg["cats"]
        top = synth.random_tree(self.n, seed)
        self.newick = synth.to_newick(top)
        self.tree = capi.RootedTree(self.newick)
        self.rates, self.freqs = synth.random_params(seed + 1 + part_seed)
        self.cat_rates = capi.gamma_cats(1.0, self.K, 0)
        self.labels = synth.tip_labels(top)
        if cfg["data"] == "iid":
            self.cols = iid_shard(self.n, off, cnt, seed + 2)
        else:
            self.cols = evolved_shard(top, self.labels, off, cnt, seed + 2 + part_seed, self.rates, self.freqs,
                                      self.cat_rates)
        self.off, self.cnt = off, cnt

    def alignment(self) -> dict:
        return {l: self.cols[i].tobytes() for i, l in enumerate(self.labels)}


def oracle_window(setup: Setup, lo: int, width: int):
    """per-site log-likelihoods of local sites [lo, lo+width) on the CPU oracle, both arithmetic modes"""
    from oracle_capi import MODE_ENGINE, MODE_REFERENCE, OraclePartition
    o = OraclePartition(setup.n, width, setup.K)
    for i, l in enumerate(setup.labels):
        o.set_tip_states(setup.tree.tip_index(l), setup.cols[i, lo:lo + width].tobytes())
    o.set_pattern_weights(np.ones(width, dtype=np.uint32))
    o.set_subst_params(setup.rates)
    o.set_frequencies(setup.freqs)
    o.set_category_rates(setup.cat_rates)
    o.set_category_weights(np.full(setup.K, 1.0 / setup.K))
    ops, pm, br = setup.tree.generate_operations(0, 0.5)
    o.update_prob_matrices(pm, br)
    o.update_clvs(ops)
    _, eng = o.root_loglikelihood(setup.tree.root_clv_index, setup.tree.root_scaler_index, persite=True, mode=MODE_ENGINE)
    _, ref = o.root_loglikelihood(setup.tree.root_clv_index, setup.tree.root_scaler_index, persite=True,
                                  mode=MODE_REFERENCE)
    o.close()
    return eng, ref


def run_config(name: str, cfg: dict, *, torch, dist, rank: int, world: int, local: int, peak_gbs: float,
               seed: int = 0x5EED0000, window: int = 256, steps: int = 3, exhaustive_branches: int = 0,
               exhaustive_iterations: int = 0, log=lambda *a: None) -> dict | None:
    from root_digger_b200 import capi
    from root_digger_b200.capi import Model, RootedTree
    from root_digger_b200.sharding import PartitionShardedModel, plan_partition_shards, plan_site_shards

    if cfg.get("sites_per_gpu"):
        cfg = dict(cfg, sites=cfg["sites"] * world)
    n, S, K, P = cfg["taxa"], cfg["sites"], cfg["cats"], cfg["partitions"]
    t_setup = time.perf_counter()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if P == 1:
        # ---- site shards, one per GPU, joined by the engine's NCCL all-reduce ------------------
        shards = plan_site_shards(S, world)
        off, cnt = shards[rank]
        setup = Setup(cfg, seed + {"cfg3": 3, "cfg5": 5}.get(name, 2), off, cnt)
        comm_id = None
        if world > 1:
            ids = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                ids = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(ids, 0)
            comm_id = bytes(ids.cpu().tolist())
        tree = RootedTree(setup.newick)
        m = Model(tree, setup.alignment(), K, site_offset=off if world > 1 else 0, global_sites=S if world > 1 else 0,
                  nranks=world, rank=rank, comm_id=comm_id)
        sharding = "sites/%d (%s sites per GPU), one NCCL all-reduce of the per-shard tree nodes per evaluation batch" % (
            world, "..".join(str(c) for c in sorted({c for _, c in shards})))
        wrap = None
        parts_local = 1
    else:
        # ---- partitions dealt to the ranks; each rank's model_t holds only its own ----------------
        owned = plan_partition_shards(P, world)[rank]
        per = S // P
        tree = None
        setups = [Setup(dict(cfg, sites=per), seed + 4, 0, per, part_seed=101 * (p + 1)) for p in owned]
        setup = setups[0] if setups else None
        if not setups:
            raise RuntimeError("cfg4 needs at most as many ranks as partitions")
        tree = RootedTree(setups[0].newick)
        cols = np.concatenate([s.cols for s in setups], axis=1)
        aln = {l: cols[i].tobytes() for i, l in enumerate(setups[0].labels)}
        ranges = [(j * per, (j + 1) * per) for j in range(len(owned))]
        m = Model(tree, aln, K, partitions=ranges if len(owned) > 1 else None)
        sharding = "partitions/%d (partition p on rank p %% %d), all-gather of the per-partition terms, summed in " \
                   "partition order" % (world, world)
        parts_local = len(owned)
        wrap = None
    m.initialize_partitions()
    if P == 1:
        m.set_params(rates=setup.rates, freqs=setup.freqs)
    else:
        for j, s in enumerate(setups):
            m.set_params(rates=s.rates, freqs=s.freqs, part=j)
        # the all-gather of the partitions' terms is issued from Python after the local call -- the path
        # measured on 8 GPUs in profiles/r02_bench_final_n8_north_star.json.  (model_t can complete the sums
        # itself, PartitionShardedModel(in_model=True): that is what search / exhaustive_search on partition
        # shards use; it is held on gloo by tests/test_sharding.py and on NCCL by tests/test_gpu_multi.py.)
        wrap = PartitionShardedModel(m, P, rank, world, dist, device="cuda", in_model=False) if world > 1 else None
    m.set_sweep_mode(m.SWEEP_DIRECTED)
    L = capi.load_engine()
    handles = [C.cast(m.L.rdh_model_partition(m.h, j), C.POINTER(capi.PartitionStruct)) for j in range(parts_local)]

    def stats():
        tot = {}
        for h in handles:
            s = capi.Stats()
            L.rdk_partition_stats(h, C.byref(s))
            for k, v in s.asdict().items():
                tot[k] = tot.get(k, 0) + v
        return tot

    def reset(timing: bool):
        for h in handles:
            L.rdk_partition_set_timing(h, 1 if timing else 0)
            L.rdk_partition_reset_stats(h)

    compute_lh = (lambda: wrap.compute_lh(0, 0.5)) if wrap else (lambda: m.compute_lh(0, 0.5))
    sweep = (lambda: wrap.sweep_root_lh()) if wrap else (lambda: m.sweep_root_lh())
    placements = m.root_count
    log(name, "set-up %.1f s" % (time.perf_counter() - t_setup))

    # ---- full evaluation ------------------------------------------------------------------------
    lh0 = compute_lh()
    compute_lh()
    reset(True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        lh = compute_lh()
    barrier()
    fe_ms = allmax((time.perf_counter() - t0) * 1e3 / steps)
    st = stats()
    fe_kernel_ms = st["program_time_ns"] * 1e-6 / max(1, st["program_timed"]) * parts_local
    fe_bytes = st["algorithmic_bytes"] / steps
    fe_gbs = fe_bytes / (fe_kernel_ms * 1e-3) / 1e9 if fe_kernel_ms > 0 else 0.0
    full_eval = {"ms": fe_ms, "kernel_ms_per_gpu": allmax(fe_kernel_ms), "algorithmic_bytes_per_gpu": fe_bytes,
                 "kernel_gbs_per_gpu": fe_gbs, "frac_of_hbm_peak": fe_gbs / peak_gbs,
                 "min_frac_over_gpus": -allmax(-fe_gbs / peak_gbs),
                 "evaluations_per_sec": 1e3 / fe_ms, "bit_reproducible": bool(lh == lh0), "logl": lh}

    # ---- the placement sweep ------------------------------------------------------------------------
    sw0 = sweep()
    reset(True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, steps - 1)):
        sw = sweep()
    barrier()
    sw_ms = allmax((time.perf_counter() - t0) * 1e3 / max(1, steps - 1))
    st = stats()
    sw_kernel_ms = st["program_time_ns"] * 1e-6 / max(1, steps - 1)
    sw_bytes = st["algorithmic_bytes"] / max(1, steps - 1)
    sw_gbs = sw_bytes / (sw_kernel_ms * 1e-3) / 1e9 if sw_kernel_ms > 0 else 0.0
    digest = hashlib.sha256(np.ascontiguousarray(sw, dtype="<f8").tobytes()).hexdigest()
    d8 = torch.tensor([int(digest[:15], 16)], dtype=torch.int64, device="cuda")
    dmax, dmin = d8.clone(), d8.clone()
    if world > 1:
        dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(dmin, op=dist.ReduceOp.MIN)
    sweep_blk = {"ms": sw_ms, "placements": int(placements), "placements_per_sec": placements / (sw_ms * 1e-3),
                 "kernel_ms_per_gpu": allmax(sw_kernel_ms), "algorithmic_bytes_per_gpu": sw_bytes,
                 "kernel_gbs_per_gpu": sw_gbs, "frac_of_hbm_peak": sw_gbs / peak_gbs,
                 "min_frac_over_gpus": -allmax(-sw_gbs / peak_gbs), "launches": st["program_launches"] // max(1, steps - 1),
                 "chunks": m.sweep_chunks, "stores_elided_per_sweep": st["stores_elided"] // max(1, steps - 1),
                 "bit_reproducible": bool(np.array_equal(sw, sw0)), "best_placement": int(np.argmax(sw)),
                 # the sweep scores root 0 at ratio 0.5 too: the reference's compute_lh == compute_lh_root invariant
                 "root0_equals_full_evaluation": bool(sw[0] == lh),
                 "digest_sha256": digest, "digest_equal_on_all_ranks": bool(dmax.item() == dmin.item())}

    # ---- a window of this rank's sites against the CPU oracle ------------------------------------------
    width = min(window, setup.cnt)
    lo = (setup.cnt // 2 // 8) * 8 if setup.cnt > width else 0
    lo = min(lo, setup.cnt - width)
    persite = np.zeros(setup.cnt)
    compute_lh()
    setup.tree.generate_operations(0, 0.5)  # root the checker's tree: the root buffers exist once it is rooted
    root_clv, root_sc = setup.tree.root_clv_index, setup.tree.root_scaler_index
    L.rdk_compute_root_loglikelihood.restype = C.c_double
    tot = L.rdk_compute_root_loglikelihood(handles[0], root_clv, root_sc, None,
                                           persite.ctypes.data_as(C.POINTER(C.c_double)))
    eng, ref = oracle_window(setup, lo, width)
    got = persite[lo:lo + width]
    rel = float(np.max(np.abs(got - ref) / np.abs(ref))) if width else 0.0
    bitwise = bool(np.array_equal(got.view(np.uint64), eng.view(np.uint64)))
    window_blk = {"sites": int(width), "local_site_offset": int(lo), "bit_identical_to_oracle_engine_mode": bitwise,
                  "max_rel_err_vs_oracle_reference_mode": rel, "within_1e-9": bool(rel <= 1e-9),
                  "all_ranks_ok": bool(allsum(0.0 if (bitwise and rel <= 1e-9) else 1.0) == 0.0),
                  "logl_after_persite_call_equals_compute_lh": bool(P > 1 or tot == lh)}

    out = {"workload": cfg["workload"], "taxa": n, "sites": S, "rate_cats": K, "partitions": P, "n_gpus": world,
           "sharding": sharding, "device_bytes_per_gpu": allmax(stats()["device_bytes"]),
           "full_evaluation": full_eval, "sweep": sweep_blk, "oracle_window": window_blk,
           "api": "librd_host.so model_t (compute_lh, sweep of suggest_roots_lh" + (", exhaustive_search)" if
                                                                                   exhaustive_branches else ")")}

    # ---- exhaustive mode on a bounded sample of branches ---------------------------------------------
    if exhaustive_branches > 0 and P == 1:
        reset(False)
        branches = max(1, min(exhaustive_branches, placements))
        num_tasks = max(1, -(-placements // branches))
        m.set_max_outer_iterations(exhaustive_iterations)
        s0 = stats()
        barrier()
        t0 = time.perf_counter()
        ids, llh, alpha = m.exhaustive_search(1e-7, 1e-7, 1e-12, 1e4, rank=0, num_tasks=num_tasks)
        barrier()
        dt = allmax(time.perf_counter() - t0)
        s1 = stats()
        evals = s1["root_evals"] - s0["root_evals"]
        full = (s1["clv_ops"] - s0["clv_ops"]) // max(1, n - 1)
        # the alpha loop: compute_dlh = 2 root-only evaluations (reference src/model.cpp:481-519)
        barrier()
        t1 = time.perf_counter()
        reps = 20
        for i in range(reps):
            m.compute_dlh(int(ids[0]), 0.3 + 0.01 * i)
        barrier()
        dlh_us = allmax((time.perf_counter() - t1) / reps * 1e6)
        out["exhaustive"] = {
            "branches": int(len(ids)), "seconds": dt, "branches_per_sec": len(ids) / dt,
            "root_evaluations": int(evals), "full_traversals": int(full), "full_evaluations_per_sec": full / dt,
            "us_per_compute_dlh": dlh_us,
            "outer_iterations_per_branch": ("capped at %d (a converged branch of this size is ~9e4 full evaluations: "
                                            "426 s in profiles/r02_bench_n8_north_star.json)" % exhaustive_iterations)
            if exhaustive_iterations else "to convergence (reference loop, <= 1000)",
            "tolerances": "atol 1e-7, pgtol 1e-7, brtol 1e-12, factor 1e4 "
                                                        "(reference src/model.cpp:1140 defaults)",
            "best_branch": int(ids[int(np.argmax(llh))]), "best_llh": float(np.max(llh)),
            "lwr_of_sample": [float(x) for x in m.lwr(llh)], "alpha_of_sample": [float(a) for a in alpha]}
    m.close()
    if rank == 0:
        return out
    return None
