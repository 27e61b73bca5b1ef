"""The C-ABI library loads here (no GPU) and exports every symbol include/rdk.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

from root_digger_b200 import _build, capi

HEADER = (_build.INCLUDE / "rdk.h").read_text()


def declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    names = re.findall(r"\b(rdk_[a-z0-9_]+)\s*\(", body)
    return sorted(set(n for n in names if n not in ("rdk_errno", "rdk_errmsg")))


def test_library_builds_loads_and_exports_every_declared_symbol():
    L = capi.load_engine()
    fns = declared_functions()
    assert len(fns) >= 30
    for name in fns:
        assert hasattr(L, name), f"{name} is declared in include/rdk.h but not exported"
    assert C.c_ulonglong.in_dll(L, "rdk_map_nt") is not None
    assert b"sm_100a" in L.rdk_version()


def test_host_library_header_is_current_and_every_declared_symbol_is_exported():
    """include/rdh.h (the host-side operator interface: rooted_tree_t, checkpoint_t, model_t, the
    optimiser components behind opaque handles) is generated from the extern "C" definitions; the
    committed text must be what the generator writes now, and both builds of the host library --
    on the CUDA engine and on the oracle -- must export every function it declares"""
    import subprocess
    import sys
    import oracle_build
    from root_digger_b200 import _build
    root = Path(__file__).resolve().parent.parent
    assert subprocess.run([sys.executable, str(root / "tools" / "gen_rdh_header.py"), "--check"]).returncode == 0, \
        "include/rdh.h is stale: run python tools/gen_rdh_header.py"
    text = re.sub(r"/\*.*?\*/", "", (root / "include" / "rdh.h").read_text(), flags=re.S)
    names = sorted(set(re.findall(r"\b(rdh_[a-z0-9_]+)\s*\(", text)))
    assert len(names) >= 85
    for path in (_build.build_host(), oracle_build.build_host_on_oracle()):
        out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True, check=True).stdout
        exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
        missing = [n for n in names if n not in exported]
        assert not missing, (path, missing)
        undeclared = sorted(e for e in exported if e.startswith("rdh_") and e not in names)
        assert not undeclared, (path, undeclared)
    # the header is plain C
    for compiler, lang in (("gcc", "c"), ("g++", "c++")):
        r = subprocess.run([compiler, "-fsyntax-only", "-I", str(root / "include"), "-x", lang, str(root / "include" / "rdh.h")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_struct_layouts_match_the_header():
    assert C.sizeof(capi.Operation) == 32                      # 8 x 4-byte fields, corax_operation_t
    assert [f[0] for f in capi.Operation._fields_] == re.findall(
        r"(?:unsigned int|int)\s+(\w+_index);", HEADER.split("typedef struct rdk_operation")[1].split("}")[0])
    assert C.sizeof(capi.Stats) == 22 * 8


def test_nucleotide_map_matches_corax_map_nt():
    L = capi.load_engine()
    m = (C.c_ulonglong * 256).in_dll(L, "rdk_map_nt")
    want = dict(A=1, C=2, G=4, T=8, U=8, R=5, Y=10, S=6, W=9, K=12, M=3, B=14, D=13, H=11, V=7, N=15, X=15, O=15)
    for ch, v in want.items():
        assert m[ord(ch)] == v and m[ord(ch.lower())] == v
    assert m[ord("-")] == 15 and m[ord("?")] == 15 and m[ord("!")] == 0 and m[ord("J")] == 0


def test_host_math_matches_oracle_bit_for_bit():
    """rdk_compute_gamma_cats is host scalar code: no GPU needed; product and oracle agree exactly"""
    import numpy as np
    from oracle_capi import gamma_cats as oracle_gc
    for alpha in (0.02, 0.2, 0.73, 1.0, 4.5, 99.0):
        for k in (1, 2, 4, 8, 16):
            for mode in (0, 1):
                a, b = capi.gamma_cats(alpha, k, mode), oracle_gc(alpha, k, mode)
                assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
    with pytest.raises(capi.EngineError):
        capi.gamma_cats(0.001, 4)


def test_sweep_chunk_hint_rule(monkeypatch):
    """rdk_sweep_chunk_hint is host arithmetic (148 SMs assumed without a device): the chunk count that
    minimises the launch planner's cost model -- (passes) x (time per instruction at the chosen elements
    per thread and warps per SM) x (1/chunks of the sweep) -- with a chunk more having to pay for
    itself; 1 once the shard fills the device in whole passes"""
    import torch
    if torch.cuda.is_available() and torch.cuda.get_device_properties(0).multi_processor_count != 148:
        pytest.skip("rule values below are for 148 SMs")
    monkeypatch.delenv("RDK_SWEEP_CHUNKS", raising=False)
    h = capi.sweep_chunk_hint
    assert h(100000, 4) == 1 and h(125000, 4) == 1 and h(500000, 4) == 1 and h(50000, 4) == 1
    assert h(62500, 4) == 2      # 1.2 passes of 4 elements per thread: two chunks waste less of the second
    assert h(25000, 4) == 2 and h(13312, 4) == 3 and h(12500, 4) == 4 and h(6250, 4) == 8
    assert h(0, 4) == 1
    for sites in (1, 10, 333, 1630, 5000, 18944, 77777):
        for k in (1, 3, 4, 16):
            assert 1 <= h(sites, k) <= 16                                 # RDK_SWEEP_MAX_CHUNKS
    monkeypatch.setenv("RDK_SWEEP_CHUNKS", "5")
    assert h(100000, 4) == 5


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.EngineError):
        capi.Partition(4, 10, 4)


def test_product_does_not_reference_the_oracle():
    pkg = _build.PKG
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.hpp")) \
            + list(pkg.rglob("*.cuh")):
        txt = path.read_text()
        assert "rd_oracle" not in txt and "rdo_" not in txt and "oracle_capi" not in txt, path
        assert "import oracle_build" not in txt and "librd_oracle" not in txt, path
