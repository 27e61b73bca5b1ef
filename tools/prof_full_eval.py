"""Small driver for ncu: a few full evaluations (+ optional placement sweep) with a given launch config.
usage: prof_full_eval.py TAXA SITES CTAS THREADS ELEMS [REPS] [sweep] [iid|evolved]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
from root_digger_b200.capi import Partition, gamma_cats, ops_array
from cases import Case, compute_lh

n, S = int(sys.argv[1]), int(sys.argv[2])
ctas, threads, elems = (int(x) for x in sys.argv[3:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
sweep = "sweep" in sys.argv[7:]
data = "evolved" if "evolved" in sys.argv[7:] else "iid"
case = Case(n, S, 4, seed=42, data=data, gamma_cats=gamma_cats)
g = Partition(n, S, 4)
case.setup(g)
g.set_launch_config(ctas, threads, elems)
sched = case.full_schedule(0, 0.5)
for _ in range(reps):
    print(compute_lh(g, sched, case.root_clv, case.root_scaler))
if sweep:
    roots = list(range(case.tree.root_count))
    pmo, mi, bl, opo, ops = case.sweep_schedule(roots, 0.5)
    out = g.sweep_root_placements(pmo, mi, bl, opo, ops_array(ops), case.root_clv, case.root_scaler)
    print("sweep best", int(np.argmax(out)), float(out.max()), g.stats())
