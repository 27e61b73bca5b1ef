"""Hardware parity of the multi-GPU paths: two processes, one GPU each, NCCL.

Site shards (the engine's own ncclAllReduce of the per-shard tree nodes, include/rdk.h
rdk_partition_attach_comm) through the C ABI and through model_t, and partitions dealt to the GPUs
(sharding.PartitionShardedModel over torch.distributed nccl) must return the bits of the same
calls on ONE GPU: full evaluation, the chunked directed sweep in one launch and cut into batches
(one collective per batch), batched root candidates, empirical frequencies, compute_dlh and
optimize_alpha (reference src/model.cpp:384-519, 679-794, 865-889); and exhaustive mode with the ROOT
PLACEMENTS dealt to the GPUs (src/model.cpp:1899-1907, a replica of all sites per GPU, no data-path
collective) must give the per-branch log-likelihoods, root positions and LWR ranking of one process
doing every branch.  Skipped below two devices."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def capi_lwr(capi, llh):
    """model_t::lwr (reference src/model.cpp:1237-1258) through the host library"""
    import ctypes as C
    L = capi.load_tree_lib()
    llh = np.ascontiguousarray(llh, dtype=np.float64)
    out = np.zeros(len(llh))
    L.rdh_model_lwr(llh.ctypes.data_as(C.POINTER(C.c_double)), len(llh), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def _devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_two_gpu_nccl_runs_return_the_bits_of_one_gpu(tmp_path):
    if _devices() < 2:
        pytest.skip("needs two GPUs")
    sys.path.insert(0, str(ROOT / "tests"))
    import gpu_multi_worker as w
    from root_digger_b200 import capi
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = tmp_path / "two_gpu.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "gpu_multi_worker.py"), str(out)]
    env = dict(os.environ)
    env.pop("RDK_SWEEP_MAX_SLOTS", None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = dict(np.load(out))

    # ---- the same calls on one GPU
    case = w.build_case()
    want = {}
    lay = case.tree.sweep_layout(2)
    g = capi.Partition(case.n, case.S, w.K, clv_buffers=lay["clv_buffers"], scale_buffers=lay["scale_buffers"],
                       prob_matrices=lay["prob_matrices"])
    case.setup(g)
    for k, v in w.engine_level(case, g, lay).items():
        want["abi_" + k] = v
    g.close()
    m = capi.Model(capi.RootedTree(case.newick), case.aln, w.K)
    m.initialize_partitions()
    m.set_params(rates=case.rates, freqs=case.freqs)
    for k, v in w.model_level(m).items():
        want["model_" + k] = v
    m.close()
    pm = capi.Model(capi.RootedTree(case.newick), case.aln, w.K, partitions=w.PARTS)
    pm.initialize_partitions()
    for p in range(len(w.PARTS)):
        pm.set_params(rates=w.part_rates(case, p), freqs=case.freqs, part=p)
    want["parts_lh"] = pm.compute_lh(4, 0.6)
    want["parts_lh_root"] = pm.compute_lh_root(4, 0.2)
    want["parts_sweep"] = pm.sweep_root_lh()
    pm.close()

    # exhaustive mode with the root ids dealt to the two GPUs == one process doing all of them
    small = w.small_case()
    em = capi.Model(capi.RootedTree(small.newick), small.aln, w.K, seed=5)
    em.initialize_partitions()
    ids, llh, alpha = em.exhaustive_search(*w.EXHAUSTIVE_TOL)
    lwr_one = em.lwr(llh)
    em.close()
    assert sorted(got["roots_ids"].astype(int).tolist()) == sorted(ids.tolist()) == list(range(2 * small.n - 3))
    order = np.argsort(got["roots_ids"])
    want["roots_ids"] = ids[np.argsort(ids)].astype(np.float64)
    want["roots_llh"], want["roots_alpha"] = llh[np.argsort(ids)], alpha[np.argsort(ids)]
    for k in ("roots_ids", "roots_llh", "roots_alpha"):
        got[k] = got[k][order]
    assert np.array_equal(np.argsort(-got["roots_llh"], kind="stable"), np.argsort(-want["roots_llh"], kind="stable"))  # LWR ranking
    assert np.array_equal(lwr_one[np.argsort(ids)].view(np.uint64), capi_lwr(capi, got["roots_llh"]).view(np.uint64))
    assert set(got) == set(want)
    for k in sorted(want):
        a = np.ascontiguousarray(np.asarray(got[k], dtype=np.float64))
        b = np.ascontiguousarray(np.asarray(want[k], dtype=np.float64))
        assert a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64)), k
    # the batched sweep is the one-launch sweep
    assert np.array_equal(got["abi_sweep"].view(np.uint64), got["abi_sweep_batched"].view(np.uint64))
