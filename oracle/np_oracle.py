"""Independent numpy/scipy restatement of the path (TEST INFRASTRUCTURE).

Used only to cross-check oracle/rd_oracle.c on small cases: P = scipy expm of
the Appendix A-2 rate matrix, Felsenstein pruning over an op list with 2^256
rescaling, site-weighted root log-likelihood (SURVEY.md Appendix A-3/A-4,
reference call sites src/model.cpp:384-413)."""
import numpy as np
from scipy.linalg import expm

NT = {c: m for c, m in zip("ACGTRYSWKMBDHVN-?XOU", [1, 2, 4, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15, 15, 15, 15, 15, 8])}


def build_q(rates, freqs):
    Q = np.zeros((4, 4))
    k = 0
    for i in range(4):
        for j in range(4):
            if i != j:
                Q[i, j] = rates[k] * freqs[j]
                k += 1
    Q[np.diag_indices(4)] = -Q.sum(1)
    return Q / -(np.asarray(freqs) * np.diag(Q)).sum()


def tip_clv(seq: bytes, K: int):
    m = np.array([NT[chr(c).upper()] for c in seq])
    v = ((m[:, None] >> np.arange(4)[None, :]) & 1).astype(float)
    return np.repeat(v[:, None, :], K, axis=1)


def loglikelihood(n_tips, tip_seqs, ops, pm, br, rates, freqs, cat_rates, cat_weights, weights, root_clv, root_scaler):
    """tip_seqs: {clv_index: bytes}; ops: list of 8-tuples in corax field order"""
    K = len(cat_rates)
    Q = build_q(rates, freqs)
    P = {int(i): np.stack([expm(Q * r * t) for r in cat_rates]) for i, t in zip(pm, br)}
    clv = {i: tip_clv(s, K) for i, s in tip_seqs.items()}
    S = len(next(iter(tip_seqs.values())))
    scal = {}
    for (p, ps, c1, m1, s1, c2, m2, s2) in ops:
        x = np.einsum("kij,skj->ski", P[m1], clv[c1])
        y = np.einsum("kij,skj->ski", P[m2], clv[c2])
        v = x * y
        cnt = (scal[s1] if s1 >= 0 else 0) + (scal[s2] if s2 >= 0 else 0) + np.zeros(S, dtype=np.int64)
        if ps >= 0:
            small = (v < 2.0 ** -256).all(axis=(1, 2))
            v[small] *= 2.0 ** 256
            cnt = cnt + small
            scal[ps] = cnt
        clv[p] = v
    term = np.einsum("k,ski,i->s", cat_weights, clv[root_clv], freqs)
    l = np.log(term)
    if root_scaler >= 0:
        l = l + scal[root_scaler] * (-256 * np.log(2.0))
    persite = l * weights
    return float(persite.sum()), persite
