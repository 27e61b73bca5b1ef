"""Shard planners (host logic of SURVEY.md section 8e).

  * sites: contiguous ranges of site patterns, one per GPU, aligned to
    RDK_SHARD_ALIGN (256) so that every shard boundary is a node boundary of
    the canonical reduction tree -> the log-likelihood does not depend on the
    number of GPUs;
  * root placements: contiguous chunks of the root-id list, the rule the
    reference uses for MPI ranks (src/model.cpp:1855-1864, 1899-1907);
  * partitions: round-robin of MSA partitions over GPUs (config 4).
"""
from __future__ import annotations

ALIGN = 256


def plan_site_shards(global_sites: int, nranks: int, align: int = ALIGN):
    """-> [(offset, count)] * nranks; offsets are multiples of `align`; counts may be 0."""
    if nranks < 1:
        raise ValueError("nranks must be positive")
    blocks = (global_sites + align - 1) // align
    base, extra = divmod(blocks, nranks)
    out, off = [], 0
    for r in range(nranks):
        nb = base + (1 if r < extra else 0)
        begin = min(off * align, global_sites)
        end = min((off + nb) * align, global_sites)
        out.append((begin, end - begin))
        off += nb
    assert sum(c for _, c in out) == global_sites
    return out


def plan_root_shards(root_ids, nranks: int):
    """contiguous chunks: beg = chunk*rank + min(mod, rank) (reference src/model.cpp:1899-1907)"""
    ids = list(root_ids)
    chunk, mod = divmod(len(ids), nranks)
    out = []
    for rank in range(nranks):
        beg = chunk * rank + min(mod, rank)
        end = chunk * (rank + 1) + min(mod, rank + 1)
        out.append(ids[beg:end])
    return out


def plan_partition_shards(n_partitions: int, nranks: int):
    return [[p for p in range(n_partitions) if p % nranks == r] for r in range(nranks)]


def replica_bytes(n_taxa: int, sites: int, rate_cats: int = 4) -> int:
    """device bytes of one partition holding `sites` patterns:
    n-1 inner CLVs of 32*K B per site, 2n-2 uint32 scale buffers, 1-byte tips, weights"""
    return sites * ((n_taxa - 1) * 32 * rate_cats + (2 * n_taxa - 2) * 4 + n_taxa + 4)


def plan_grid(nranks: int, n_taxa: int, sites: int, rate_cats: int = 4, budget_bytes: float = 150e9,
              force: str = "auto"):
    """2-D decomposition: nranks = site_groups x root_groups.

    Root placements shard with no data-path collective but need a replica of all the
    sites they score; sites shard with one all-reduce per evaluation.  So: as few site
    shards as HBM allows (the smallest divisor G_s of nranks whose shard fits
    `budget_bytes`), and the remaining factor distributes the root placements
    (exhaustive mode, src/model.cpp:1899-1907).  cfg2 (6.4 GB) and cfg3 (137 GB) fit a
    replica -> (1, N); cfg5 (1.35 TB) -> (8, 1).
    force = "sites" -> (N, 1); "roots" -> (1, N).
    Rank r works on site shard r % G_s for root chunk r // G_s."""
    if nranks < 1:
        raise ValueError("nranks must be positive")
    if force == "sites":
        return nranks, 1
    if force == "roots":
        return 1, nranks
    for gs in range(1, nranks + 1):
        if nranks % gs:
            continue
        shard_sites = max(c for _, c in plan_site_shards(sites, gs))
        if replica_bytes(n_taxa, shard_sites, rate_cats) <= budget_bytes:
            return gs, nranks // gs
    return nranks, 1


def _exchange_fn_type():
    """the C type of model_t::partition_exchange_fn, created once: every instance must hand ctypes the
    same class it declared in argtypes"""
    import ctypes as C
    global _EXCHANGE_FN
    try:
        return _EXCHANGE_FN
    except NameError:
        _EXCHANGE_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_size_t, C.c_size_t, C.POINTER(C.c_double),
                                   C.c_void_p)
        return _EXCHANGE_FN


class PartitionShardedModel:
    """BASELINE cfg4: the partitions of a multi-partition alignment are dealt out to
    the ranks (`plan_partition_shards`: partition p lives on rank p % nranks); every rank holds a
    `capi.Model` of ITS partitions only.  Per-partition parameter optimisation needs no
    communication at all (the reference's BFGS closures are per partition, src/model.cpp:1544-1547);
    every log-likelihood the reference sums over partitions (src/model.cpp:397,429) is completed
    INSIDE model_t through its partition exchange (model_t::set_partition_exchange): one all-gather
    of the per-partition terms per evaluation batch, after which every rank adds them in global
    partition order -- the order a single process uses.  So not only `compute_lh` /
    `compute_lh_root` / `sweep_root_lh` but `compute_dlh`, `optimize_alpha`, `search` and
    `exhaustive_search` run on the shards, take the same decisions on every rank and return the bits
    of a single process holding all partitions.

    `dist` is torch.distributed (initialised; nccl with `device="cuda"` or gloo with "cpu").
    Every rank must make the same calls in the same order (each one is a collective)."""

    def __init__(self, local_model, n_partitions: int, rank: int, nranks: int, dist, device="cpu",
                 in_model: bool = True):
        """in_model=False keeps model_t unaware of the other ranks: only compute_lh, compute_lh_root
        and sweep_root_lh are available, completed by an all-gather issued from Python after the
        local call (same collective, same ordered sum, same bits)."""
        import ctypes as C
        self.in_model = bool(in_model)
        self.m, self.P, self.rank, self.nranks, self.dist, self.device = local_model, n_partitions, rank, nranks, dist, device
        if nranks > n_partitions:
            # a rank without a partition would have no model_t to call (and nothing to all-gather)
            raise ValueError("partition sharding needs at most as many ranks (%d) as partitions (%d)"
                             % (nranks, n_partitions))
        self.owned = plan_partition_shards(n_partitions, nranks)
        if local_model.partition_count != len(self.owned[rank]):
            raise ValueError("the local model must hold exactly this rank's partitions")
        self.slots = max(len(o) for o in self.owned)
        self.exchanges = 0
        self._error = None
        self._exchange = None
        if not self.in_model:
            return
        fn_t = _exchange_fn_type()

        def exchange(local_ptr, n_local, count, all_ptr, _user):
            import numpy as np
            out = np.ctypeslib.as_array(all_ptr, shape=(self.P, count))
            try:
                local = np.ctypeslib.as_array(local_ptr, shape=(n_local, count))
                out[:] = self._gather_terms(local)
                self.exchanges += 1
            except Exception as e:  # nothing may propagate through the C++ frames: poison the sums instead
                self._error = e
                out[:] = float("nan")

        self._exchange = fn_t(exchange)  # kept alive as long as the model uses it
        L = local_model.L
        L.rdh_model_set_partition_exchange.argtypes = [C.c_void_p, C.POINTER(C.c_uint), C.c_uint, C.c_uint, fn_t,
                                                       C.c_void_p]
        L.rdh_model_rng_state.argtypes = [C.c_void_p]
        L.rdh_model_rng_state.restype = C.c_ulonglong
        L.rdh_model_set_rng_state.argtypes = [C.c_void_p, C.c_ulonglong]
        L.rdh_model_set_rng_state.restype = None
        L.rdh_model_discard_rng.argtypes = [C.c_void_p, C.c_ulonglong]
        L.rdh_model_discard_rng.restype = None
        L.rdh_model_first_partition_without_empirical_freqs.argtypes = [C.c_void_p]
        mine = (C.c_uint * len(self.owned[rank]))(*self.owned[rank])
        if not L.rdh_model_set_partition_exchange(local_model.h, mine, len(self.owned[rank]), n_partitions,
                                                  self._exchange, None):
            raise RuntimeError(L.rdh_last_error().decode())

    def close(self):
        import ctypes as C
        if self.m is not None and self.m.h and self._exchange is not None:
            self.m.L.rdh_model_set_partition_exchange(self.m.h, None, 0, 0, C.cast(None, _exchange_fn_type()), None)
        self.m = None

    def _gather_terms(self, local):
        """local: [local partitions][n] -> [all partitions][n], every rank"""
        import numpy as np
        import torch
        local = np.atleast_2d(np.asarray(local, dtype=np.float64))
        n = local.shape[1]
        buf = torch.zeros((self.slots, n), dtype=torch.float64, device=self.device)
        buf[:local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(self.device)
        out = torch.zeros((self.nranks, self.slots, n), dtype=torch.float64, device=self.device)
        self.dist.all_gather_into_tensor(out.view(-1), buf.view(-1)) if self.device != "cpu" else \
            self.dist.all_gather(list(out.unbind(0)), buf)
        allv = out.cpu().numpy()
        terms = np.zeros((self.P, n))
        for r, parts in enumerate(self.owned):
            for j, p in enumerate(parts):
                terms[p] = allv[r, j]
        return terms

    def _all_ranks(self, value: int):
        """[value of rank 0, value of rank 1, ...] on every rank"""
        import torch
        mine = torch.tensor([int(value)], dtype=torch.int64, device=self.device)
        out = [torch.zeros_like(mine) for _ in range(self.nranks)]
        self.dist.all_gather(out, mine)
        return [int(t.item()) for t in out]

    def _call(self, fn, *a, **kw):
        if not self.in_model:
            raise RuntimeError("this call needs the sums over partitions inside model_t (in_model=True)")
        try:
            return fn(*a, **kw)
        finally:
            if self._error is not None:
                e, self._error = self._error, None
                raise RuntimeError("partition exchange failed: %r" % (e,))

    @staticmethod
    def _ordered_sum(terms):
        import numpy as np
        total = np.zeros(terms.shape[1])
        for p in range(terms.shape[0]):  # left to right over partitions, like model_t::compute_lh
            total = total + terms[p]
        return total

    def initialize_partitions(self, uniform_freqs: bool = False):
        if not self.in_model:
            return self.m.initialize_partitions(uniform_freqs=uniform_freqs)
        return self._initialize_partitions_in_step(uniform_freqs)

    def _initialize_partitions_in_step(self, uniform_freqs: bool):
        """model_t::initialize_partitions[_uniform_freqs] on every rank.  A partition whose empirical
        frequencies have a zero entry makes a single process throw when it gets there; here every
        rank raises (not only the one that holds it) and every generator is left where the single
        process's would be, so that the caller can fall back to uniform frequencies on all ranks as
        RootDigger's main does."""
        L, h = self.m.L, self.m.h
        if not uniform_freqs:
            state = L.rdh_model_rng_state(h)
            bad = L.rdh_model_first_partition_without_empirical_freqs(h)
            if bad == -2:
                raise RuntimeError(L.rdh_last_error().decode())
            first = [self.owned[r][b] for r, b in enumerate(self._all_ranks(bad)) if b >= 0]
            if first:
                L.rdh_model_set_rng_state(h, state)
                L.rdh_model_discard_rng(h, min(first))  # the draws of the partitions before the failing one
                raise RuntimeError("One of the state frequenices is zero while using emperical frequencies")
        self.m.initialize_partitions(uniform_freqs=uniform_freqs)

    def set_params(self, global_partition: int, **kw):
        """parameters of one GLOBAL partition (a no-op on the ranks that do not hold it)"""
        if global_partition in self.owned[self.rank]:
            self.m.set_params(part=self.owned[self.rank].index(global_partition), **kw)

    # ---- everything below is the local model_t: its sums over partitions are global ----------
    def compute_lh(self, rid: int, ratio: float = 0.5) -> float:
        if not self.in_model:
            self.m.compute_lh(rid, ratio)
            return float(self._ordered_sum(self._gather_terms(self.m.last_partition_lh()[:, None]))[0])
        return self._call(self.m.compute_lh, rid, ratio)

    def compute_lh_root(self, rid: int, ratio: float = 0.5) -> float:
        if not self.in_model:
            self.m.compute_lh_root(rid, ratio)
            return float(self._ordered_sum(self._gather_terms(self.m.last_partition_lh()[:, None]))[0])
        return self._call(self.m.compute_lh_root, rid, ratio)

    def compute_dlh(self, rid: int, ratio: float = 0.5):
        return self._call(self.m.compute_dlh, rid, ratio)

    def move_root(self, rid: int, ratio: float = 0.5):
        return self.m.move_root(rid, ratio)

    def optimize_alpha(self, rid: int, ratio: float = 0.5, atol: float = 1e-7) -> float:
        return self._call(self.m.optimize_alpha, rid, ratio, atol)

    def optimize_root_location(self, min_roots: int = 1, root_ratio: float = 0.05):
        return self._call(self.m.optimize_root_location, min_roots, root_ratio)

    def sweep_root_lh(self):
        if not self.in_model:
            self.m.sweep_root_lh()
            return self._ordered_sum(self._gather_terms(self.m.last_sweep_partition_lh()))
        return self._call(self.m.sweep_root_lh)

    def search(self, *a, **kw):
        """model_t::search over ALL partitions; every rank runs every start (rank / num_tasks of the
        local model stay 0 / 1: the partitions are what is distributed, not the starts) and logs
        complete records (the parameters of all partitions)"""
        return self._call(self.m.search, *a, **kw)

    def exhaustive_search(self, *a, **kw):
        return self._call(self.m.exhaustive_search, *a, **kw)

    def set_checkpoint(self, prefix):
        """one log per rank ("<prefix>.part<rank>of<nranks>.ckp"), never a shared file (two ranks
        would both append every record).  Each holds COMPLETE records -- model_t gathers the fitted
        parameters of all partitions through the exchange -- so any of them is a checkpoint of the
        whole run in the reference's format."""
        if prefix is not None:
            prefix = "%s.part%dof%d" % (prefix, self.rank, self.nranks)
        return self.m.set_checkpoint(prefix)

    def __getattr__(self, name):
        # settings and accessors without a sum over partitions (set_max_outer_iterations, set_checkpoint,
        # set_batched_probes, root_count, lwr, ...) are the local model's
        if name in ("m", "in_model"):
            raise AttributeError(name)
        return getattr(self.m, name)

    def last_partition_lh(self):
        """the terms of ALL partitions of the last compute_lh / compute_lh_root, every rank"""
        return self._gather_terms(self.m.last_partition_lh()[:, None])[:, 0]
