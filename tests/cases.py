"""Shared parity-test driver: one seeded case, identical call sequences for the
oracle partition and the CUDA engine partition (same names as the reference's
model_t steps: compute_lh, compute_lh_root, move_root, suggest_roots_lh)."""
from __future__ import annotations

import numpy as np

from root_digger_b200 import synth
from root_digger_b200.capi import RootedTree


class Case:
    def __init__(self, n_taxa: int, sites: int, K: int, seed: int, data: str = "evolved", alpha: float = 0.7,
                 weights: str = "ones", mean_brlen: float = 0.05, gamma_cats=None, tree_lib=None):
        self.n, self.S, self.K, self.seed = n_taxa, sites, K, seed
        top = synth.random_tree(n_taxa, seed, mean_brlen)
        self.newick = synth.to_newick(top)
        self.tree = RootedTree(self.newick, lib=tree_lib)
        self.rates, self.freqs = synth.random_params(seed + 1)
        if gamma_cats is None:
            from oracle_capi import gamma_cats as gc
            gamma_cats = gc
        self.cat_rates = gamma_cats(alpha, K, 0)
        self.cat_weights = np.full(K, 1.0 / K)
        if data == "evolved":
            self.aln = synth.simulate_alignment(top, sites, seed + 2, self.rates, self.freqs, self.cat_rates)
        else:
            self.aln = synth.iid_alignment(synth.tip_labels(top), sites, seed + 2)
        rng = np.random.default_rng(seed + 3)
        if weights == "ones":
            self.weights = np.ones(sites, dtype=np.uint32)
        else:
            self.weights = rng.integers(1, 5, sites).astype(np.uint32)
        if data == "ambiguous":
            # sprinkle IUPAC codes and gaps
            codes = np.frombuffer(b"ACGTRYSWKMBDHVN-", dtype=np.uint8)
            for l in list(self.aln):
                a = np.frombuffer(self.aln[l], dtype=np.uint8).copy()
                m = rng.random(sites) < 0.15
                a[m] = codes[rng.integers(0, len(codes), int(m.sum()))]
                self.aln[l] = a.tobytes()

    # ---- set-up (model_t::initialize_partitions, reference src/model.cpp:1297-1306)
    def setup(self, part, site_slice: slice | None = None):
        for label, seq in self.aln.items():
            s = seq if site_slice is None else seq[site_slice]
            part.set_tip_states(self.tree.tip_index(label), s)
        w = self.weights if site_slice is None else self.weights[site_slice]
        part.set_pattern_weights(w)
        part.set_subst_params(self.rates)
        part.set_frequencies(self.freqs)
        part.set_category_rates(self.cat_rates)
        part.set_category_weights(self.cat_weights)

    # ---- schedules (generated once, applied to every backend)
    def full_schedule(self, rid: int, ratio: float = 0.5):
        return self.tree.generate_operations(rid, ratio)

    def derivative_schedule(self, rid: int, ratio: float):
        return self.tree.generate_derivative_operations(rid, ratio)

    def move_schedule(self, rid: int, ratio: float = 0.5):
        return self.tree.generate_root_update_operations(rid, ratio)

    @property
    def root_clv(self):
        return self.tree.root_clv_index

    @property
    def root_scaler(self):
        return self.tree.root_scaler_index

    def sweep_schedule(self, roots, ratio: float = 0.5):
        """what suggest_roots_lh (reference src/model.cpp:865-889) issues per root:
        move_root (path ops + their matrices) then compute_lh_root (2 matrices + root op)"""
        pm_off, op_off, mi, bl, ops = [0], [0], [], [], []
        for rid in roots:
            o, m, b = self.tree.generate_root_update_operations(rid, ratio)
            mi.extend(m.tolist())
            bl.extend(b.tolist())
            ops.extend(o)
            op1, m1, b1 = self.tree.generate_derivative_operations(rid, ratio)
            mi.extend(m1.tolist())
            bl.extend(b1.tolist())
            ops.append(op1)
            pm_off.append(len(mi))
            op_off.append(len(ops))
        return (np.array(pm_off, dtype=np.uint32), np.array(mi, dtype=np.uint32), np.array(bl),
                np.array(op_off, dtype=np.uint32), ops)


def compute_lh(part, sched, root_clv, root_scaler, **kw):
    """model_t::compute_lh (reference src/model.cpp:384-413)"""
    ops, pm, br = sched
    part.update_prob_matrices(pm, br)
    part.update_clvs(ops)
    return part.root_loglikelihood(root_clv, root_scaler, **kw)


def compute_lh_root(part, dsched, root_clv, root_scaler, **kw):
    """model_t::compute_lh_root (reference src/model.cpp:415-452)"""
    op, pm, br = dsched
    part.update_prob_matrices(pm, br)
    part.update_clvs([op])
    return part.root_loglikelihood(root_clv, root_scaler, **kw)


def move_root(part, msched):
    """model_t::move_root (reference src/model.cpp:823-854)"""
    ops, pm, br = msched
    if len(pm):
        part.update_prob_matrices(pm, br)
    if len(ops):
        part.update_clvs(ops)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same_bits(a, b) -> bool:
    return np.array_equal(bits(np.asarray(a, dtype=np.float64)), bits(np.asarray(b, dtype=np.float64)))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(d / den)) if d.size else 0.0
