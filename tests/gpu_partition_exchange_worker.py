"""One rank of the 2-GPU run of tests/test_gpu_zz_partition_exchange.py: three partitions dealt to two
GPUs, the sums over partitions completed INSIDE model_t (model_t::set_partition_exchange, all-gather
over torch.distributed nccl).  `whole_model_results` is also run by the test on one GPU holding all
three partitions and must give the same bits."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import gpu_multi_worker as base  # noqa: E402  (the case and the partition layout of the NCCL parity run)

SEARCH_ARGS = dict(min_roots=2, root_ratio=0.02, atol=1e-2, pgtol=1e-2, brtol=1e-3, factor=1e13)


def whole_model_results(m):
    out = {}
    m.compute_lh(0, 0.5)
    probes = []
    for rid, x in ((2, 0.3), (5, 1.0), (7, 0.0)):
        m.compute_lh(rid, 0.5)
        probes += list(m.compute_dlh(rid, x))
        probes.append(m.optimize_alpha(rid, 0.5, 1e-9))
    out["probes"] = np.array(probes)
    out["sweep"] = np.asarray(m.sweep_root_lh())
    m.set_max_outer_iterations(2)
    rid, alpha, lh = m.search(strategy="random", **SEARCH_ARGS)
    out["search"] = np.array([float(rid), alpha, lh])
    return out


def build_model(case, parts_global, mine):
    from root_digger_b200 import capi
    cols = {l: b"".join(s[parts_global[p][0]:parts_global[p][1]] for p in mine) for l, s in case.aln.items()}
    ranges, pos = [], 0
    for p in mine:
        ranges.append((pos, pos + parts_global[p][1] - parts_global[p][0]))
        pos = ranges[-1][1]
    return capi.Model(capi.RootedTree(case.newick), cols, base.K, seed=17, early_stop=True, partitions=ranges)


def main(out_path):
    import torch
    import torch.distributed as dist
    from root_digger_b200 import sharding
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    case = base.build_case()
    mine = sharding.plan_partition_shards(len(base.PARTS), world)[rank]
    m = build_model(case, base.PARTS, mine)
    sm = sharding.PartitionShardedModel(m, len(base.PARTS), rank, world, dist, device="cuda")
    sm.initialize_partitions()
    res = whole_model_results(sm)
    res["exchanges"] = np.array([float(sm.exchanges)])
    np.savez(out_path + ".rank%d.npz" % rank, **res)
    dist.barrier()
    sm.close()
    m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
