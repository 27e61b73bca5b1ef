"""Exploratory timing of the placement sweep under different launch configs."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, torch
from root_digger_b200.capi import Partition, gamma_cats, ops_array
from cases import Case, compute_lh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
case = Case(n, S, 4, seed=42, data="iid", gamma_cats=gamma_cats)
g = Partition(n, S, 4)
g.set_stream(torch.cuda.current_stream().cuda_stream)
case.setup(g)
roots = list(range(case.tree.root_count))
compute_lh(g, case.full_schedule(0, .5), case.root_clv, case.root_scaler)
sw = case.sweep_schedule(roots, 0.5)
pmo, mi, bl, opo, ops = sw
arr = ops_array(ops)
print("placements", len(roots), "ops", len(ops), "pmats", len(mi))
for ctas, threads, elems in [(6,128,1),(4,192,1),(3,128,2),(4,96,2),(2,128,4),(4,64,4)]:
    g.set_launch_config(ctas, threads, elems)
    for rep in range(2):
        compute_lh(g, case.full_schedule(0, .5), case.root_clv, case.root_scaler)
        case.tree.root_by(0, .5)
        g.reset_stats()
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        t0 = time.time(); a.record()
        out = g.sweep_root_placements(pmo, mi, bl, opo, arr, case.root_clv, case.root_scaler)
        b.record(); torch.cuda.synchronize(); dt = time.time() - t0
    st = g.stats()
    print(f"cfg {ctas},{threads},{elems}: wall {dt*1e3:.1f} ms, device {a.elapsed_time(b):.1f} ms -> {len(roots)/dt:.0f} placements/s, alg {st['algorithmic_bytes']/1e9:.1f} GB -> {st['algorithmic_bytes']/a.elapsed_time(b)/1e6:.0f} GB/s, launches {st['kernel_launches']}, best lh {out.max():.3f}", flush=True)
