"""Hardware run of the in-model partition exchange (BASELINE cfg4): three partitions
dealt to two GPUs, every sum over partitions completed inside model_t through an all-gather over
torch.distributed nccl (sharding.PartitionShardedModel, model_t::set_partition_exchange) --
compute_dlh, optimize_alpha, the placement sweep and a search from shuffled starts must return, on
both ranks, the bits of ONE GPU holding all three partitions.  tests/test_sharding.py holds the same
on gloo + the oracle.  Sorted last among the GPU tests; skipped below two devices."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def test_partition_sharded_model_over_nccl_returns_the_bits_of_one_gpu(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sys.path.insert(0, str(ROOT / "tests"))
    import gpu_partition_exchange_worker as w
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "parts")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "gpu_partition_exchange_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]

    case = w.base.build_case()
    m = w.build_model(case, w.base.PARTS, list(range(len(w.base.PARTS))))
    m.initialize_partitions()
    want = w.whole_model_results(m)
    m.close()
    for rank in (0, 1):
        got = dict(np.load(out + ".rank%d.npz" % rank))
        assert got.pop("exchanges")[0] > 10
        assert set(got) == set(want)
        for k in sorted(want):
            a, b = np.ascontiguousarray(got[k], dtype=np.float64), np.ascontiguousarray(want[k], dtype=np.float64)
            assert a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64)), (rank, k, a, b)
