"""The reference's bundled test fixtures (copied verbatim from
/root/reference/test/data into tests/golden/ref_fixtures) as parity-test cases."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from root_digger_b200.capi import RootedTree

FX = Path(__file__).resolve().parent / "golden" / "ref_fixtures"
FILES = {"single": ("single.phy", "single.tree"), "10.fasta": ("10.fasta", "10.tree"),
         "101.phy": ("101.phy", "101.tree")}


def read_alignment(path: Path) -> dict:
    txt = path.read_text()
    if txt.lstrip().startswith(">"):
        d, cur = {}, None
        for line in txt.splitlines():
            line = line.strip()
            if line.startswith(">"):
                cur = line[1:].strip()
                d[cur] = ""
            elif line:
                d[cur] += line
        return d
    lines = [l for l in txt.splitlines() if l.strip()]
    n, length = (int(x) for x in lines[0].split()[:2])
    d = {}
    for l in lines[1:1 + n]:
        parts = l.split()
        d[parts[0]] = "".join(parts[1:])
    assert all(len(s) == length for s in d.values()), "only one-line-per-taxon PHYLIP fixtures are bundled"
    return d


def compress(aln: dict):
    """site pattern compression (identical columns merged, weights added), sorted pattern order"""
    labels = list(aln)
    M = np.array([np.frombuffer(aln[l].upper().encode(), dtype=np.uint8) for l in labels])
    cols, inv, counts = np.unique(M.T, axis=0, return_inverse=True, return_counts=True)
    out = {l: cols[:, i].tobytes() for i, l in enumerate(labels)}
    return out, counts.astype(np.uint32)


def load(name: str):
    a, t = FILES[name]
    return {"name": name, "alignment": read_alignment(FX / a), "tree_path": FX / t}


class FixtureCase:
    """Same interface as cases.Case, on a bundled fixture; parameters follow the reference's
    own test set-up: rates from test/src/model.cpp:16, uniform frequencies (:19-21)."""

    RATES = np.array([.34, .42, .24, .74, .16, .88, .75, .54, .20, .06, .08, .41])

    def __init__(self, fx, K: int, tree_lib=None, alpha: float = 1.0):
        from oracle_capi import gamma_cats
        self.tree = RootedTree(path=str(fx["tree_path"]), lib=tree_lib)
        self.aln, self.weights = compress(fx["alignment"])
        self.n = self.tree.tip_count
        self.S = len(self.weights)
        self.K = K
        self.rates = self.RATES.copy()
        self.freqs = np.full(4, 0.25)
        self.cat_rates = gamma_cats(alpha, K, 0)
        self.cat_weights = np.full(K, 1.0 / K)

    def setup(self, part, site_slice=None):
        for label, seq in self.aln.items():
            s = seq if site_slice is None else seq[site_slice]
            part.set_tip_states(self.tree.tip_index(label), s)
        part.set_pattern_weights(self.weights if site_slice is None else self.weights[site_slice])
        part.set_subst_params(self.rates)
        part.set_frequencies(self.freqs)
        part.set_category_rates(self.cat_rates)
        part.set_category_weights(self.cat_weights)

    def full_schedule(self, rid, ratio=0.5):
        return self.tree.generate_operations(rid, ratio)

    def derivative_schedule(self, rid, ratio):
        return self.tree.generate_derivative_operations(rid, ratio)

    def move_schedule(self, rid, ratio=0.5):
        return self.tree.generate_root_update_operations(rid, ratio)

    root_clv = property(lambda s: s.tree.root_clv_index)
    root_scaler = property(lambda s: s.tree.root_scaler_index)
