"""The drop-in boundary proven with the reference's OWN callers.

RootDigger's src/model.cpp, src/tree.cpp, src/msa.cpp, src/checkpoint.cpp and src/util.cpp are
compiled UNCHANGED -- from where they lie in the reference checkout, never copied -- against
root_digger_b200/compat/corax/corax.h, the header that serves the coraxlib calls they make
(src/model.cpp:159-168 ... 466, src/tree.cpp:12-620, src/msa.cpp:18-88) from the engine's C ABI
(oracle/oracle_build.py build_reference_sources).  Through tests/ref_build/ref_capi.cpp they run the
call sequence of src/main.cpp:513-640 on the bundled fixtures; the engine host's model_t mirror
(root_digger_b200/host) must return the same bits: chosen root branch, root position, final
log-likelihood, every per-branch likelihood, LWR ranking.

Here (no GPU) both sides run on the oracle behind the same ABI; tests/test_gpu_model.py repeats the
search on the CUDA engine.  The built library travels to the GPU box; without the reference checkout
and without a built library the tests skip."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import fixtures
import oracle_build
import oracle_capi
from root_digger_b200 import capi

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint)


class ReferenceBuild:
    """ctypes face of tests/ref_build/ref_capi.cpp"""

    def __init__(self, backend: str):
        path = oracle_build.build_reference_sources(backend)
        if path is None or not path.exists():
            pytest.skip("the reference checkout is not here and no library was built earlier")
        self.L = C.CDLL(str(path))
        L = self.L
        L.rdref_last_error.restype = C.c_char_p
        L.rdref_create.restype = C.c_void_p
        L.rdref_create.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_ulonglong, C.c_int, C.c_char_p]
        L.rdref_destroy.argtypes = [C.c_void_p]
        L.rdref_root_count.argtypes = [C.c_void_p]
        L.rdref_root_count.restype = C.c_uint
        L.rdref_search.argtypes = [C.c_void_p, C.c_uint, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                   C.c_int, _up, _dp, _dp]
        L.rdref_exhaustive.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, _up, _dp, _dp,
                                       C.c_uint, _up, _up, _dp]
        L.rdref_all_root_lh.argtypes = [C.c_void_p, _dp, C.c_uint]

    def model(self, fx_name: str, rate_cats: int, seed: int, early_stop: bool, tmp):
        a, t = fixtures.FILES[fx_name]
        h = self.L.rdref_create(str(fixtures.FX / t).encode(), str(fixtures.FX / a).encode(), rate_cats, seed,
                                1 if early_stop else 0, os.path.join(tmp, "ref").encode())
        if not h:
            raise RuntimeError(self.L.rdref_last_error().decode())
        return C.c_void_p(h)

    def check(self, rc):
        if not rc:
            raise RuntimeError(self.L.rdref_last_error().decode())


@pytest.fixture(scope="module")
def ref():
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    return ReferenceBuild("oracle")


@pytest.fixture(scope="module")
def mirror_lib():
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    return capi.load_tree_lib(oracle_build.build_host_on_oracle())


def mirror_model(lib, fx_name, rate_cats, seed, early_stop):
    a, t = fixtures.FILES[fx_name]
    tree = capi.RootedTree(path=str(fixtures.FX / t), lib=lib)
    m = capi.Model.from_files(tree, fixtures.FX / a, None, rate_cats, seed=seed, early_stop=early_stop)
    try:
        m.initialize_partitions(uniform_freqs=False)
    except RuntimeError:
        m.initialize_partitions(uniform_freqs=True)
    m.L.rdh_model_initialize(m.h) if hasattr(m.L, "rdh_model_initialize") else None
    return m


def same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def check_search(ref, mirror_lib, fx_name, K, strategy):
    with tempfile.TemporaryDirectory() as tmp:
        h = ref.model(fx_name, K, 7, True, tmp)
        rid, alpha, lh = C.c_uint(), C.c_double(), C.c_double()
        ref.check(ref.L.rdref_search(h, 2, 0.05, 1e-3, 1e-3, 1e-4, 1e12, strategy, C.byref(rid), C.byref(alpha),
                                     C.byref(lh)))
        n_roots = ref.L.rdref_root_count(h)
        ref.L.rdref_destroy(h)
    m = mirror_model(mirror_lib, fx_name, K, 7, True)
    assert m.root_count == n_roots
    got = m.search(2, 0.05, 1e-3, 1e-3, 1e-4, 1e12, strategy=["random", "midpoint", "modified_mad"][strategy])
    m.close()
    assert got[0] == rid.value
    assert same([got[1], got[2]], [alpha.value, lh.value]), (got, rid.value, alpha.value, lh.value)


def check_exhaustive(ref, mirror_lib, K):
    with tempfile.TemporaryDirectory() as tmp:
        h = ref.model("10.fasta", K, 3, False, tmp)
        n = ref.L.rdref_root_count(h)
        ids, llh, alpha = np.zeros(n, dtype=np.uint32), np.zeros(n), np.zeros(n)
        got_n, best_id, best_lh = C.c_uint(), C.c_uint(), C.c_double()
        ref.check(ref.L.rdref_exhaustive(h, 1e-2, 1e-2, 1e-3, 1e13, ids.ctypes.data_as(_up), llh.ctypes.data_as(_dp),
                                         alpha.ctypes.data_as(_dp), n, C.byref(got_n), C.byref(best_id),
                                         C.byref(best_lh)))
        ref.L.rdref_destroy(h)
    assert got_n.value == n
    m = mirror_model(mirror_lib, "10.fasta", K, 3, False)
    mids, mllh, malpha = m.exhaustive_search(1e-2, 1e-2, 1e-3, 1e13)
    order_ref, order_m = np.argsort(ids), np.argsort(mids)
    assert np.array_equal(ids[order_ref], mids[order_m])
    assert same(llh[order_ref], mllh[order_m]) and same(alpha[order_ref], malpha[order_m])
    assert int(mids[np.argmax(mllh)]) == best_id.value
    assert same(m.lwr(mllh[order_m]), m.lwr(llh[order_ref]))
    m.close()


def check_every_root(ref, mirror_lib):
    with tempfile.TemporaryDirectory() as tmp:
        h = ref.model("101.phy", 4, 11, True, tmp)
        n = ref.L.rdref_root_count(h)
        out = np.zeros(n)
        ref.check(ref.L.rdref_all_root_lh(h, out.ctypes.data_as(_dp), n))
        ref.L.rdref_destroy(h)
    m = mirror_model(mirror_lib, "101.phy", 4, 11, True)
    want = m.compute_all_root_lh()
    m.close()
    assert same(out, want)


# ---- the CPU suite keeps the light cases (both sides on the oracle); tests/test_gpu_reference_sources.py
# ---- runs the heavy ones (4 categories, exhaustive mode) on the CUDA engine
@pytest.mark.parametrize("K,strategy", [(1, 1), (2, 0)])
def test_search_with_the_reference_sources_equals_the_mirror(ref, mirror_lib, K, strategy):
    """model_t::search (src/model.cpp:1008-1138) from the reference's own file: same root, same
    position on the branch, same log-likelihood, bit for bit"""
    check_search(ref, mirror_lib, "10.fasta", K, strategy)


def test_every_root_with_the_reference_sources_equals_the_mirror(ref, mirror_lib):
    """compute_all_root_lh (src/model.cpp:1737-1746): move_root + compute_lh per root through the
    reference's own tree.cpp schedules and the compat tree module"""
    check_every_root(ref, mirror_lib)
