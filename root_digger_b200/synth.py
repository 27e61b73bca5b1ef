"""Seeded synthetic inputs of the benchmark shapes.

  * random_tree: random unrooted binary topology by sequential addition,
    branch lengths Exp(mean 0.05) clamped to [1e-4, 1.0]
  * random_params: 12 UNREST rates U(1e-4, 1) (the distribution of
    random_params, reference src/model.cpp:87-93), pi from a Dirichlet
  * simulate_alignment: DNA evolved down the tree under UNREST + discrete Gamma
    (states ACGT only), or iid-uniform columns

Only data generation lives here; none of it is on the measured path.
"""
from __future__ import annotations

import numpy as np

DNA = np.frombuffer(b"ACGT", dtype=np.uint8)


class Node:
    __slots__ = ("kids", "label", "length")

    def __init__(self, label=None, length=0.0):
        self.kids = []
        self.label = label
        self.length = length


def random_tree(n_taxa: int, seed: int, mean_brlen: float = 0.05) -> Node:
    """Unrooted binary tree as a top node with three children."""
    assert n_taxa >= 3
    rng = np.random.default_rng(seed)

    def brlen():
        return float(min(1.0, max(1e-4, rng.exponential(mean_brlen))))

    top = Node()
    edges = []  # (parent, index in parent.kids)
    for i in range(3):
        top.kids.append(Node(f"t{i}", brlen()))
        edges.append((top, i))
    for t in range(3, n_taxa):
        parent, idx = edges[int(rng.integers(len(edges)))]
        old = parent.kids[idx]
        mid = Node(None, old.length * 0.5)
        old.length = max(1e-4, old.length * 0.5)
        mid.length = max(1e-4, mid.length)
        new = Node(f"t{t}", brlen())
        mid.kids = [old, new]
        parent.kids[idx] = mid
        edges.append((mid, 0))
        edges.append((mid, 1))
    return top


def to_newick(node: Node) -> str:
    """newick text with three children at the top level (an unrooted tree)"""
    return _newick_iter(node) + ";"


def _newick_iter(root: Node) -> str:
    parts = []
    stack = [("open", root)]
    while stack:
        kind, nd = stack.pop()
        if kind == "text":
            parts.append(nd)
            continue
        if not nd.kids:
            parts.append(f"{nd.label}:{nd.length!r}")
            continue
        parts.append("(")
        tail = ")" if nd is root else f"){nd.label or ''}:{nd.length!r}"
        stack.append(("text", tail))
        for i, k in enumerate(reversed(nd.kids)):
            stack.append(("open", k))
            if i != len(nd.kids) - 1:
                stack.append(("text", ","))
    return "".join(parts)


def random_params(seed: int):
    rng = np.random.default_rng(seed)
    rates = rng.uniform(1e-4, 1.0, 12)
    freqs = rng.dirichlet(np.full(4, 20.0))
    return rates, freqs


def build_q(rates, freqs) -> np.ndarray:
    """Q_ij = r_ij * pi_j, unit mean rate."""
    Q = np.zeros((4, 4))
    k = 0
    for i in range(4):
        for j in range(4):
            if i != j:
                Q[i, j] = rates[k] * freqs[j]
                k += 1
    Q[np.diag_indices(4)] = -Q.sum(1)
    mu = -(freqs * np.diag(Q)).sum()
    return Q / mu


def _expm(A: np.ndarray) -> np.ndarray:
    # plain scaling and squaring with a Taylor core: data generation only
    n = max(0, int(np.ceil(np.log2(max(np.abs(A).sum(1).max(), 1e-300)))) + 4)
    B = A / (2.0 ** n)
    E = np.eye(4)
    term = np.eye(4)
    for i in range(1, 18):
        term = term @ B / i
        E = E + term
    for _ in range(n):
        E = E @ E
    return E


def simulate_alignment(top: Node, sites: int, seed: int, rates, freqs, cat_rates) -> dict:
    """Evolve `sites` columns down the tree; returns {label: bytes}."""
    rng = np.random.default_rng(seed)
    Q = build_q(rates, freqs)
    K = len(cat_rates)
    cats = rng.integers(0, K, sites)
    root_states = rng.choice(4, size=sites, p=np.asarray(freqs) / np.sum(freqs)).astype(np.int8)
    out = {}
    stack = [(top, root_states)]
    while stack:
        nd, st = stack.pop()
        if not nd.kids:
            out[nd.label] = DNA[st].tobytes()
            continue
        for k in nd.kids:
            child = np.empty(sites, dtype=np.int8)
            u = rng.random(sites)
            for c in range(K):
                sel = cats == c
                if not sel.any():
                    continue
                P = np.clip(_expm(Q * cat_rates[c] * k.length), 0.0, None)
                cum = np.cumsum(P / P.sum(1, keepdims=True), axis=1)
                cs = cum[st[sel]]
                child[sel] = (u[sel, None] > cs[:, :3]).sum(1)
            stack.append((k, child))
    return out


def iid_alignment(labels, sites: int, seed: int) -> dict:
    rng = np.random.default_rng(seed)
    return {l: DNA[rng.integers(0, 4, sites)].tobytes() for l in labels}


def tip_labels(top: Node):
    out = []
    stack = [top]
    while stack:
        nd = stack.pop()
        if not nd.kids:
            out.append(nd.label)
        else:
            stack.extend(nd.kids)
    return sorted(out, key=lambda s: int(s[1:]))
