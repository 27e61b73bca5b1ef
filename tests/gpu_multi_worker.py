"""One rank of the 2-GPU NCCL parity run (tests/test_gpu_multi.py launches two of these with
torch.distributed.run).  Every quantity below is also computed on ONE GPU by the test and must have
the same bits: the reduction tree is defined over the global site index and the all-reduced tree
nodes receive +0.0 from the shards that do not own them."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

N_TAXA, SITES, K, SEED = 60, 5000, 4, 4242
PARTS = [(0, 1700), (1700, 2600), (2600, 5000)]


def build_case():
    from cases import Case
    from root_digger_b200.capi import gamma_cats
    return Case(N_TAXA, SITES, K, seed=SEED, data="ambiguous", weights="ones", gamma_cats=gamma_cats)


EXHAUSTIVE_TOL = (1e-2, 1e-2, 1e-3, 1e13)


def small_case():
    from cases import Case
    from root_digger_b200.capi import gamma_cats
    return Case(9, 700, K, seed=SEED + 9, data="evolved", weights="ones", gamma_cats=gamma_cats)


def part_rates(case, p):
    return np.roll(case.rates, p) * (1.0 + 0.07 * p)


def engine_level(case, g, lay, monkey_slots=None):
    """compute_lh + chunked directed sweep (+ the same sweep cut into batches) + batched root candidates"""
    from cases import compute_lh
    from root_digger_b200 import capi
    out = {}
    out["lh"] = compute_lh(g, case.full_schedule(3, 0.4), case.root_clv, case.root_scaler)
    *csw, cpos, coff = case.tree.generate_chunked_sweep_operations(layout=lay)
    flags = capi.RDK_SWEEP_KEEP_ROOT | capi.RDK_SWEEP_DISCARD
    sw = np.empty(len(cpos))
    sw[cpos] = g.sweep_root_placements(*csw, case.root_clv, case.root_scaler, flags=flags, chunk_offsets=coff)
    out["sweep"] = sw
    os.environ["RDK_SWEEP_MAX_SLOTS"] = "16"
    sw2 = np.empty(len(cpos))
    sw2[cpos] = g.sweep_root_placements(*csw, case.root_clv, case.root_scaler, flags=flags, chunk_offsets=coff)
    del os.environ["RDK_SWEEP_MAX_SLOTS"]
    out["sweep_batched"] = sw2
    op, pm, br = case.derivative_schedule(3, 0.4)
    total = float(br[0] + br[1])
    cand = np.array([[total * a, total * (1 - a)] for a in (0.0, 0.25, 0.4, 1.0)]).ravel()
    out["multi"] = g.root_loglikelihood_multi(op, cand)
    out["freqs"] = g.empirical_frequencies()
    return out


def model_level(m):
    out = {"lh": m.compute_lh(2, 0.3), "sweep": m.sweep_root_lh(), "lh_root": m.compute_lh_root(2, 0.8)}
    lh, dlh = m.compute_dlh(2, 0.35)
    out["dlh"] = np.array([lh, dlh])
    m.move_root(7, 0.5)
    out["alpha"] = m.optimize_alpha(7, 0.5, 1e-9)
    return out


def main():
    import torch
    import torch.distributed as dist
    from root_digger_b200 import capi, sharding
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def comm_id():
        ids = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ids = torch.tensor(list(capi.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(ids, 0)
        return bytes(ids.cpu().tolist())

    case = build_case()
    off, cnt = sharding.plan_site_shards(SITES, world)[rank]
    res = {}
    # ---- the C ABI with a site shard per GPU
    lay = case.tree.sweep_layout(2)
    g = capi.Partition(case.n, cnt, K, device=local, clv_buffers=lay["clv_buffers"], scale_buffers=lay["scale_buffers"],
                       prob_matrices=lay["prob_matrices"])
    case.setup(g, slice(off, off + cnt))
    g.set_shard(off, SITES)
    g.attach_comm(world, rank, comm_id())
    for k, v in engine_level(case, g, lay).items():
        res["abi_" + k] = v
    g.close()
    # ---- model_t with a site shard per GPU
    aln = {l: s[off:off + cnt] for l, s in case.aln.items()}
    m = capi.Model(capi.RootedTree(case.newick), aln, K, site_offset=off, global_sites=SITES, nranks=world, rank=rank,
                   comm_id=comm_id())
    m.initialize_partitions()
    m.set_params(rates=case.rates, freqs=case.freqs)
    for k, v in model_level(m).items():
        res["model_" + k] = v
    m.close()
    # ---- partitions dealt to the GPUs (BASELINE cfg4 in miniature)
    mine = sharding.plan_partition_shards(len(PARTS), world)[rank]
    cols = {l: b"".join(s[PARTS[p][0]:PARTS[p][1]] for p in mine) for l, s in case.aln.items()}
    ranges, pos = [], 0
    for p in mine:
        ranges.append((pos, pos + PARTS[p][1] - PARTS[p][0]))
        pos = ranges[-1][1]
    pm = capi.Model(capi.RootedTree(case.newick), cols, K, partitions=ranges)
    pm.initialize_partitions()
    for j, p in enumerate(mine):
        pm.set_params(rates=part_rates(case, p), freqs=case.freqs, part=j)
    sm = sharding.PartitionShardedModel(pm, len(PARTS), rank, world, dist, device="cuda", in_model=False)
    res["parts_lh"] = sm.compute_lh(4, 0.6)
    res["parts_lh_root"] = sm.compute_lh_root(4, 0.2)
    res["parts_sweep"] = sm.sweep_root_lh()
    pm.close()
    # ---- root placements dealt to the GPUs (exhaustive mode, reference src/model.cpp:1899-1907): every
    # ---- rank holds a replica of all sites and optimises its contiguous chunk of the root ids
    small = small_case()
    em = capi.Model(capi.RootedTree(small.newick), small.aln, K, seed=5)
    em.initialize_partitions()
    ids, llh, alpha = em.exhaustive_search(*EXHAUSTIVE_TOL, rank=rank, num_tasks=world)
    em.close()
    mine = torch.full((3, 64), float("nan"), dtype=torch.float64, device="cuda")
    mine[0, :len(ids)] = torch.from_numpy(ids.astype(np.float64))
    mine[1, :len(ids)] = torch.from_numpy(llh)
    mine[2, :len(ids)] = torch.from_numpy(alpha)
    allv = torch.zeros((world, 3, 64), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(allv.view(-1), mine.view(-1))
    allv = allv.cpu().numpy()
    keep = ~np.isnan(allv[:, 0, :])
    res["roots_ids"] = np.concatenate([allv[r, 0, keep[r]] for r in range(world)])
    res["roots_llh"] = np.concatenate([allv[r, 1, keep[r]] for r in range(world)])
    res["roots_alpha"] = np.concatenate([allv[r, 2, keep[r]] for r in range(world)])
    if rank == 0:
        np.savez(sys.argv[1], **{k: np.asarray(v) for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
