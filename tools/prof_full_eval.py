"""Small driver for ncu: a few full evaluations (+ optional sweep) with a given launch config."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np
from root_digger_b200.capi import Partition, gamma_cats
from cases import Case, compute_lh

n, S = int(sys.argv[1]), int(sys.argv[2])
ctas, threads, elems = (int(x) for x in sys.argv[3:6])
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
case = Case(n, S, 4, seed=42, data="iid", gamma_cats=gamma_cats)
g = Partition(n, S, 4)
case.setup(g)
g.set_launch_config(ctas, threads, elems)
sched = case.full_schedule(0, 0.5)
for _ in range(reps):
    print(compute_lh(g, sched, case.root_clv, case.root_scaler))
