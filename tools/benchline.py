"""bench.py output (files given as arguments, else stdin) -> one summary line per JSON line"""
import sys, json, fileinput
for l in fileinput.input():
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        fe = d.get("full_evaluation") or {}
        print(round(d["value"]), "pl/s", round(d["ms_per_step"], 3), "ms/step  frac", round(d["roofline"]["frac"], 3),
              " full-eval kernel ms", round(fe.get("kernel_ms", 0), 3), " launch ms", round(d["roofline"]["avg_launch_ms"], 3),
              " chunks", d["config"].get("sweep_chunks"), " logl", d["logl_root0"])
    elif l:
        print(l[:200])
