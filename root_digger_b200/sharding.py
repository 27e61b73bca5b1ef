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


class PartitionShardedModel:
    """BASELINE cfg4: the partitions of a multi-partition alignment are dealt out to
    the ranks (`plan_partition_shards`: partition p lives on rank p % nranks); every rank holds a
    `capi.Model` of ITS partitions only.  Per-partition parameter optimisation needs no
    communication at all (the reference's BFGS closures are per partition, src/model.cpp:1544-1547);
    `compute_lh` / `compute_lh_root` / `sweep_root_lh` need the SUM over all partitions
    (src/model.cpp:397,429): one all-gather of the per-partition terms per call, after which every
    rank adds them in global partition order -- the order a single process uses -- so the result
    has the same bits on 1, 2 or 8 ranks.

    `dist` is torch.distributed (initialised; nccl with `device="cuda"` or gloo with "cpu")."""

    def __init__(self, local_model, n_partitions: int, rank: int, nranks: int, dist, device="cpu"):
        self.m, self.P, self.rank, self.nranks, self.dist, self.device = local_model, n_partitions, rank, nranks, dist, device
        if nranks > n_partitions:
            # a rank without a partition would have no model_t to call (and nothing to all-gather)
            raise ValueError("partition sharding needs at most as many ranks (%d) as partitions (%d)"
                             % (nranks, n_partitions))
        self.owned = plan_partition_shards(n_partitions, nranks)
        if local_model.partition_count != len(self.owned[rank]):
            raise ValueError("the local model must hold exactly this rank's partitions")
        self.slots = max(len(o) for o in self.owned)

    def _gather_terms(self, local):
        """local: [local partitions][n] -> [all partitions][n], every rank"""
        import numpy as np
        import torch
        local = np.atleast_2d(np.asarray(local, dtype=np.float64))
        n = local.shape[1]
        buf = torch.zeros((self.slots, n), dtype=torch.float64, device=self.device)
        buf[:local.shape[0]] = torch.from_numpy(local).to(self.device)
        out = torch.zeros((self.nranks, self.slots, n), dtype=torch.float64, device=self.device)
        self.dist.all_gather_into_tensor(out.view(-1), buf.view(-1)) if self.device != "cpu" else \
            self.dist.all_gather(list(out.unbind(0)), buf)
        allv = out.cpu().numpy()
        terms = np.zeros((self.P, n))
        for r, parts in enumerate(self.owned):
            for j, p in enumerate(parts):
                terms[p] = allv[r, j]
        return terms

    @staticmethod
    def _ordered_sum(terms):
        import numpy as np
        total = np.zeros(terms.shape[1])
        for p in range(terms.shape[0]):  # left to right over partitions, like model_t::compute_lh
            total = total + terms[p]
        return total

    def compute_lh(self, rid: int, ratio: float = 0.5) -> float:
        self.m.compute_lh(rid, ratio)
        return float(self._ordered_sum(self._gather_terms(self.m.last_partition_lh()[:, None]))[0])

    def compute_lh_root(self, rid: int, ratio: float = 0.5) -> float:
        self.m.compute_lh_root(rid, ratio)
        return float(self._ordered_sum(self._gather_terms(self.m.last_partition_lh()[:, None]))[0])

    def sweep_root_lh(self):
        self.m.sweep_root_lh()
        return self._ordered_sum(self._gather_terms(self.m.last_sweep_partition_lh()))
