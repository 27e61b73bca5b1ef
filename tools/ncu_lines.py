"""Per-source-line instruction counts and stall samples from an .ncu-rep (needs -lineinfo).
usage: ncu_lines.py REPORT.ncu-rep [kernel-index] [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 1; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; out = []; kernel = 0; infile = False
for r in rows:
    if r and r[0] == "File Path":
        infile = r[1].endswith("rdk_kernels.cuh")
        if infile: kernel += 1
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and infile and r and r[0] not in ("", "File Path", "Function Name") and kernel == which:
        col = {h: i for i, h in enumerate(hdr)}
        try:
            st = {h[6:]: int(r[i]) for h, i in col.items() if h.startswith("stall_") and "Not Issued" not in h and r[i] not in ("", "-")}
            out.append((int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]), int(r[0]), r[1].strip()[:90], st))
        except Exception:
            pass
tot_i = sum(o[1] for o in out); tot_s = sum(o[0] for o in out)
print("total instrs", tot_i, "samples", tot_s)
for s, n, l, src, st in sorted(out, reverse=True)[:top]:
    tops = ", ".join("%s %d" % (k, v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
    print("%5.1f%% samp %5.1f%% instr  L%-4d %-90s | %s" % (100 * s / tot_s, 100 * n / tot_i, l, src, tops))
