"""Generates tests/golden/golden_v1.json from the CPU oracle (run HERE, in the
build container; the JSON is committed and travels to the GPU box).

The reference's own tests pin no numeric log-likelihood, so these
vectors freeze the oracle's outputs on the reference's bundled fixtures
(test/data: single, 10.fasta, 101.phy -- copied under ref_fixtures/) and on
seeded synthetic cases.  Doubles are stored as hex strings (bit exact).

usage: python tests/golden/make_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import fixtures  # noqa: E402
from cases import Case, compute_lh, compute_lh_root  # noqa: E402
from oracle_capi import MODE_ENGINE, MODE_REFERENCE, OraclePartition, gamma_cats, load_oracle  # noqa: E402


def hx(x):
    return float(x).hex()


def main():
    out = {"version": 1, "fixtures": {}, "synthetic": [], "gamma_cats": [], "expm": []}
    L = load_oracle()
    # gamma categories
    for alpha in (0.2, 0.5, 1.0, 2.5, 50.0):
        for k in (2, 4, 8):
            for mode in (0, 1):
                out["gamma_cats"].append({"alpha": alpha, "k": k, "mode": mode,
                                          "rates": [hx(v) for v in gamma_cats(alpha, k, mode)]})
    # expm of the UNREST rate matrix
    rng = np.random.default_rng(2024)
    import ctypes as C
    dp = C.POINTER(C.c_double)
    for i in range(12):
        r = rng.uniform(1e-4, 1, 12)
        pi = rng.dirichlet(np.ones(4) * 5)
        t = float(np.exp(rng.uniform(-10, 3))) if i else 0.0
        Q = np.zeros(16)
        E = np.zeros(16)
        L.rdo_build_q_nonrev(r.ctypes.data_as(dp), pi.ctypes.data_as(dp), Q.ctypes.data_as(dp))
        A = Q * t
        L.rdo_expm4(A.ctypes.data_as(dp), E.ctypes.data_as(dp))
        out["expm"].append({"rates": [hx(v) for v in r], "freqs": [hx(v) for v in pi], "t": hx(t),
                            "Q": [hx(v) for v in Q], "P": [hx(v) for v in E]})
    # reference fixtures: log-likelihood of every root, K = 1 and K = 4
    for name in ("single", "10.fasta", "101.phy"):
        for K in (1, 4):
            fx = fixtures.load(name)
            case = fixtures.FixtureCase(fx, K)
            o = OraclePartition(case.n, case.S, K)
            case.setup(o)
            rec = {"K": K, "patterns": case.S, "taxa": case.n, "roots": case.tree.root_count,
                   "freqs": [hx(v) for v in case.freqs], "rates": [hx(v) for v in case.rates], "lh": []}
            step = 1 if case.tree.root_count <= 20 else 13
            for rid in range(0, case.tree.root_count, step):
                for ratio in (0.5, 0.2):
                    a = compute_lh(o, case.full_schedule(rid, ratio), case.root_clv, case.root_scaler,
                                   mode=MODE_REFERENCE)
                    b = o.root_loglikelihood(case.root_clv, case.root_scaler, mode=MODE_ENGINE)
                    rec["lh"].append({"root": rid, "ratio": ratio, "reference": hx(a), "engine": hx(b)})
            out["fixtures"][f"{name}:K{K}"] = rec
    # synthetic cases (the same generator the GPU tests use)
    for (n, S, K, data, weights, seed) in [(5, 7, 4, "evolved", "random", 11), (33, 300, 2, "ambiguous", "random", 12),
                                           (300, 64, 4, "iid", "ones", 13), (24, 100, 8, "evolved", "ones", 14)]:
        case = Case(n, S, K, seed=seed, data=data, weights=weights)
        o = OraclePartition(n, S, K)
        case.setup(o)
        a, ps = compute_lh(o, case.full_schedule(1, 0.3), case.root_clv, case.root_scaler, persite=True,
                           mode=MODE_REFERENCE)
        b = o.root_loglikelihood(case.root_clv, case.root_scaler, mode=MODE_ENGINE)
        c = compute_lh_root(o, case.derivative_schedule(1, 0.7), case.root_clv, case.root_scaler, mode=MODE_ENGINE)
        out["synthetic"].append({"n": n, "S": S, "K": K, "data": data, "weights": weights, "seed": seed,
                                 "reference": hx(a), "engine": hx(b), "engine_ratio_0.7": hx(c),
                                 "persite_head": [hx(v) for v in ps[:8]],
                                 "max_scaler": int(o.get_scaler(case.root_scaler).max())})
    path = Path(__file__).with_name("golden_v1.json")
    path.write_text(json.dumps(out, indent=1))
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
