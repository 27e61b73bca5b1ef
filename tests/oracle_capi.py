"""ctypes binding of the CPU ORACLE (oracle/librd_oracle.so) -- test infrastructure.

Same method names as root_digger_b200.capi.Partition so that parity tests drive
the oracle and the CUDA engine with identical call sequences.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

import sys

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "oracle"))
import oracle_build  # noqa: E402
from root_digger_b200.capi import Operation, ops_array

MODE_REFERENCE = 0
MODE_ENGINE = 1


class OraclePartitionStruct(C.Structure):
    _fields_ = [
        ("tips", C.c_uint), ("clv_buffers", C.c_uint), ("states", C.c_uint), ("sites", C.c_uint),
        ("rate_matrices", C.c_uint), ("prob_matrices", C.c_uint), ("rate_cats", C.c_uint),
        ("scale_buffers", C.c_uint), ("attributes", C.c_uint),
        ("clv", C.POINTER(C.POINTER(C.c_double))),
        ("pmatrix", C.POINTER(C.POINTER(C.c_double))),
        ("scale_buffer", C.POINTER(C.POINTER(C.c_uint))),
        ("subst_params", C.POINTER(C.POINTER(C.c_double))),
        ("frequencies", C.POINTER(C.POINTER(C.c_double))),
        ("rates", C.POINTER(C.c_double)),
        ("rate_weights", C.POINTER(C.c_double)),
        ("prop_invar", C.POINTER(C.c_double)),
        ("pattern_weights", C.POINTER(C.c_uint)),
        ("invariant", C.POINTER(C.c_int)),
    ]


_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint)
_pp = C.POINTER(OraclePartitionStruct)
_lib = None


def _ptr(a, t):
    return a.ctypes.data_as(t)


def load_oracle() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = oracle_build.ORACLE / "librd_oracle.so"
    if not path.exists():
        oracle_build.build_oracle()
    L = C.CDLL(str(path))
    L.rdo_partition_create.restype = _pp
    L.rdo_partition_create.argtypes = [C.c_uint] * 9
    L.rdo_partition_destroy.argtypes = [_pp]
    L.rdo_partition_destroy.restype = None
    L.rdo_set_tip_states.argtypes = [_pp, C.c_uint, C.c_void_p, C.c_char_p]
    L.rdo_set_pattern_weights.argtypes = [_pp, _up]
    L.rdo_set_subst_params.argtypes = [_pp, C.c_uint, _dp]
    L.rdo_set_frequencies.argtypes = [_pp, C.c_uint, _dp]
    L.rdo_set_category_rates.argtypes = [_pp, _dp]
    L.rdo_set_category_weights.argtypes = [_pp, _dp]
    L.rdo_update_prob_matrices.argtypes = [_pp, _up, _up, _dp, C.c_uint]
    L.rdo_update_clvs.argtypes = [_pp, C.POINTER(Operation), C.c_uint]
    L.rdo_update_clvs.restype = None
    L.rdo_update_clvs_mt.argtypes = [_pp, C.POINTER(Operation), C.c_uint, C.c_int]
    L.rdo_update_clvs_mt.restype = None
    L.rdo_compute_root_loglikelihood_mode.argtypes = [_pp, C.c_uint, C.c_int, _up, _dp, C.c_int]
    L.rdo_compute_root_loglikelihood_mode.restype = C.c_double
    L.rdo_compute_root_loglikelihood_mt.argtypes = [_pp, C.c_uint, C.c_int, C.c_int]
    L.rdo_compute_root_loglikelihood_mt.restype = C.c_double
    L.rdo_compute_gamma_cats.argtypes = [C.c_double, C.c_uint, _dp, C.c_int]
    L.rdo_msa_empirical_frequencies.argtypes = [_pp]
    L.rdo_msa_empirical_frequencies.restype = C.c_void_p
    L.rdo_build_q_nonrev.argtypes = [_dp, _dp, _dp]
    L.rdo_build_q_nonrev.restype = None
    L.rdo_set_q_convention.argtypes = [C.c_int]
    L.rdo_set_q_convention.restype = None
    L.rdo_get_q_convention.restype = C.c_int
    L.rdo_expm4.argtypes = [_dp, _dp]
    L.rdo_expm4.restype = None
    L.rdo_log.argtypes = [C.c_double]
    L.rdo_log.restype = C.c_double
    L.rdo_pairwise_sum.argtypes = [_dp, C.c_ulong]
    L.rdo_pairwise_sum.restype = C.c_double
    L.rdo_set_default_mode.argtypes = [C.c_int]
    L.rdo_set_default_mode.restype = None
    _lib = L
    return L


def gamma_cats(alpha: float, k: int, mode: int = 0) -> np.ndarray:
    L = load_oracle()
    out = np.zeros(k)
    assert L.rdo_compute_gamma_cats(alpha, k, _ptr(out, _dp), mode) == 1
    return out


class OraclePartition:
    def __init__(self, tips: int, sites: int, rate_cats: int = 4, *, clv_buffers=None, prob_matrices=None,
                 scale_buffers=None, attributes: int = (1 << 11) | (1 << 10)):
        self.L = load_oracle()
        branches = 2 * tips - 2
        self.tips, self.sites, self.K = tips, sites, rate_cats
        self.clv_buffers = branches if clv_buffers is None else clv_buffers
        self.prob_matrices = branches if prob_matrices is None else prob_matrices
        self.scale_buffers = branches if scale_buffers is None else scale_buffers
        self.p = self.L.rdo_partition_create(tips, self.clv_buffers, 4, sites, 1, self.prob_matrices, rate_cats,
                                             self.scale_buffers, attributes)
        assert self.p
        self._zeros = (C.c_uint * max(1, rate_cats))()

    def close(self):
        if getattr(self, "p", None):
            self.L.rdo_partition_destroy(self.p)
            self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tip_states(self, tip: int, seq: bytes):
        rc = self.L.rdo_set_tip_states(self.p, tip, C.addressof(C.c_ulonglong.in_dll(self.L, "rdo_map_nt")), seq)
        if rc != 1:
            raise ValueError(C.c_char_p.in_dll(self.L, "rdo_errmsg").value if False else "illegal state")

    def set_pattern_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.uint32)
        self.L.rdo_set_pattern_weights(self.p, _ptr(w, _up))

    def set_subst_params(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        self.L.rdo_set_subst_params(self.p, 0, _ptr(r, _dp))

    def set_frequencies(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        self.L.rdo_set_frequencies(self.p, 0, _ptr(f, _dp))

    def set_category_rates(self, r):
        r = np.ascontiguousarray(r, dtype=np.float64)
        self.L.rdo_set_category_rates(self.p, _ptr(r, _dp))

    def set_category_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64)
        self.L.rdo_set_category_weights(self.p, _ptr(w, _dp))

    def update_prob_matrices(self, matrix_indices, branch_lengths):
        mi = np.ascontiguousarray(matrix_indices, dtype=np.uint32)
        bl = np.ascontiguousarray(branch_lengths, dtype=np.float64)
        rc = self.L.rdo_update_prob_matrices(self.p, self._zeros, _ptr(mi, _up), _ptr(bl, _dp), len(mi))
        if rc != 1:
            raise ValueError("rdo_update_prob_matrices failed")

    def update_clvs(self, ops, threads: int = 0):
        arr = ops if isinstance(ops, C.Array) else ops_array(ops)
        if threads > 1:
            self.L.rdo_update_clvs_mt(self.p, arr, len(ops), threads)
        else:
            self.L.rdo_update_clvs(self.p, arr, len(ops))

    def root_loglikelihood(self, clv_index: int, scaler_index: int, persite: bool = False, mode: int = MODE_REFERENCE):
        ps = np.zeros(self.sites) if persite else None
        v = self.L.rdo_compute_root_loglikelihood_mode(self.p, clv_index, scaler_index, self._zeros,
                                                       _ptr(ps, _dp) if persite else None, mode)
        return (v, ps) if persite else v

    def root_loglikelihood_mt(self, clv_index: int, scaler_index: int, threads: int):
        return self.L.rdo_compute_root_loglikelihood_mt(self.p, clv_index, scaler_index, threads)

    def root_loglikelihood_multi(self, root_op, branch_length_pairs, mode: int = MODE_REFERENCE):
        """restates rdk_root_loglikelihood_multi with the three reference calls; restores state"""
        op = root_op if isinstance(root_op, Operation) else Operation(*root_op)
        bl = np.asarray(branch_length_pairs, dtype=np.float64).reshape(-1, 2)
        saved_clv = self.get_clv(op.parent_clv_index).copy()
        saved_sc = self.get_scaler(op.parent_scaler_index).copy() if op.parent_scaler_index >= 0 else None
        saved_p = [self.get_pmatrix(op.child1_matrix_index).copy(), self.get_pmatrix(op.child2_matrix_index).copy()]
        out = []
        for t1, t2 in bl:
            self.update_prob_matrices([op.child1_matrix_index, op.child2_matrix_index], [t1, t2])
            self.update_clvs([op])
            out.append(self.root_loglikelihood(op.parent_clv_index, op.parent_scaler_index, mode=mode))
        n = self.sites * self.K * 4
        C.memmove(self.p.contents.clv[op.parent_clv_index], saved_clv.ctypes.data, n * 8)
        if saved_sc is not None:
            C.memmove(self.p.contents.scale_buffer[op.parent_scaler_index], saved_sc.ctypes.data, self.sites * 4)
        for mi, sp in zip((op.child1_matrix_index, op.child2_matrix_index), saved_p):
            C.memmove(self.p.contents.pmatrix[mi], sp.ctypes.data, self.K * 16 * 8)
        return np.array(out)

    def sweep_root_placements(self, pm_offsets, matrix_indices, branch_lengths, op_offsets, ops,
                              root_clv_index: int, root_scaler_index: int, mode: int = MODE_REFERENCE):
        out = []
        for q in range(len(pm_offsets) - 1):
            a, b = pm_offsets[q], pm_offsets[q + 1]
            if b > a:
                self.update_prob_matrices(matrix_indices[a:b], branch_lengths[a:b])
            a, b = op_offsets[q], op_offsets[q + 1]
            if b > a:
                self.update_clvs(list(ops[a:b]))
            out.append(self.root_loglikelihood(root_clv_index, root_scaler_index, mode=mode))
        return np.array(out)

    def empirical_frequencies(self) -> np.ndarray:
        ptr = self.L.rdo_msa_empirical_frequencies(self.p)
        out = np.array(C.cast(ptr, _dp)[0:4])
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        libc.free(ptr)
        return out

    def get_clv(self, idx: int) -> np.ndarray:
        n = self.sites * self.K * 4
        return np.ctypeslib.as_array(self.p.contents.clv[idx], shape=(n,)).reshape(self.sites, self.K, 4)

    def get_scaler(self, idx: int) -> np.ndarray:
        return np.ctypeslib.as_array(self.p.contents.scale_buffer[idx], shape=(self.sites,))

    def get_pmatrix(self, idx: int) -> np.ndarray:
        return np.ctypeslib.as_array(self.p.contents.pmatrix[idx], shape=(self.K * 16,)).reshape(self.K, 4, 4)
