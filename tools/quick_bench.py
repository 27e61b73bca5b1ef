"""Exploratory timing of the program kernel under different launch configs (not the bench contract)."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import numpy as np, torch
from root_digger_b200.capi import Partition, gamma_cats
from cases import Case, compute_lh

n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
K = 4
t0 = time.time()
case = Case(n, S, K, seed=42, data="iid", gamma_cats=gamma_cats)
print("case built %.1fs" % (time.time() - t0), flush=True)
g = Partition(n, S, K)
g.set_stream(torch.cuda.current_stream().cuda_stream)
t0 = time.time(); case.setup(g); g.sync(); print("setup %.1fs" % (time.time() - t0), flush=True)
sched = case.full_schedule(0, 0.5)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))
res = []
for ctas, threads, elems in [(6,128,1),(4,192,1),(3,256,1),(5,160,1),(3,128,2),(4,96,2),(6,64,2),(2,128,4),(4,64,4),(3,96,4)]:
    g.set_launch_config(ctas, threads, elems)
    g.reset_stats()
    lh = compute_lh(g, sched, case.root_clv, case.root_scaler)
    st = g.stats()
    tmin, tmed = timed(lambda: compute_lh(g, sched, case.root_clv, case.root_scaler))
    gbs = st["algorithmic_bytes"] / tmin / 1e6
    print(f"cfg ctas={ctas} thr={threads} E={elems}: full eval min {tmin:.3f} ms med {tmed:.3f} ms  alg {st['algorithmic_bytes']/1e9:.2f} GB -> {gbs:.0f} GB/s  lh={lh:.6f} launches={st['kernel_launches']}", flush=True)
    res.append((tmin, ctas, threads, elems))
best = min(res); print("best", best)
g.set_launch_config(*best[1:])
roots = list(range(case.tree.root_count))
compute_lh(g, case.full_schedule(0, .5), case.root_clv, case.root_scaler)
sw = case.sweep_schedule(roots, 0.5)
g.reset_stats()
t0 = time.time(); out = g.sweep_root_placements(*sw, case.root_clv, case.root_scaler); dt = time.time() - t0
st = g.stats()
print(f"sweep {len(roots)} placements: {dt*1e3:.2f} ms wall -> {len(roots)/dt:.0f} placements/s; ops {st['clv_ops']} alg {st['algorithmic_bytes']/1e9:.2f} GB -> {st['algorithmic_bytes']/dt/1e9:.0f} GB/s launches {st['kernel_launches']}", flush=True)
# sequential ABI sweep
from cases import move_root, compute_lh_root
compute_lh(g, case.full_schedule(0, .5), case.root_clv, case.root_scaler)
scheds = []
for rid in roots[:200]:
    scheds.append((case.move_schedule(rid, .5), case.derivative_schedule(rid, .5)))
compute_lh(g, case.full_schedule(0, .5), case.root_clv, case.root_scaler)
# note: schedules were generated along the same sequence of roots
case.tree.root_by(0, .5)
t0 = time.time()
for ms, ds in scheds:
    move_root(g, ms); compute_lh_root(g, ds, case.root_clv, case.root_scaler)
dt = time.time() - t0
print(f"sequential ABI sweep (python overhead included): {len(scheds)/dt:.0f} placements/s")
