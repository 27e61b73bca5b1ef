"""Shard planners (host logic of SURVEY.md section 8e).

  * sites: contiguous ranges of site patterns, one per GPU, aligned to
    RDK_SHARD_ALIGN (1024) so that every shard boundary is a node boundary of
    the canonical reduction tree -> the log-likelihood does not depend on the
    number of GPUs;
  * root placements: contiguous chunks of the root-id list, the rule the
    reference uses for MPI ranks (src/model.cpp:1855-1864, 1899-1907);
  * partitions: round-robin of MSA partitions over GPUs (config 4).
"""
from __future__ import annotations

ALIGN = 1024


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
