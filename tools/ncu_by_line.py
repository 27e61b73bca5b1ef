"""Executed instructions / stall samples per CUDA source line for one launch of an .ncu-rep.
usage: ncu_by_line.py REPORT.ncu-rep OBJ_OR_LIB KERNEL_SUBSTRING LAUNCH_INDEX [top]"""
import csv, collections, io, subprocess, sys, os
rep, obj, pat, which = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
here = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[col["Instructions Executed"]].isdigit()]
lines = [l.rstrip("\n").split("\t") for l in
         subprocess.run([sys.executable, os.path.join(here, "sass_lines.py"), obj, pat], capture_output=True, text=True).stdout.split("\n") if l]
n = len(lines)
slices = [data[i:i + n] for i in range(0, len(data), n)]
print("launches in report:", len(slices), "sass instructions per kernel:", n)
sl = slices[which]
agg = collections.Counter(); smp = collections.Counter(); cnt = collections.Counter()
src_text = {}
for r, l in zip(sl, lines):
    x = int(r[col["Instructions Executed"]]); agg[l[1]] += x; smp[l[1]] += int(r[col["# Samples"]]); cnt[l[1]] += 1
tot = sum(agg.values()); ts = sum(smp.values())
print("launch %d: %.3fG warp instructions, %d samples" % (which, tot / 1e9, ts))
srcfile = {}
for k, v in agg.most_common(top):
    f, ln = k.rsplit(":", 1)
    text = ""
    for cand in (os.path.join(here, "..", "root_digger_b200", "csrc", f),):
        if os.path.exists(cand):
            if cand not in srcfile: srcfile[cand] = open(cand).read().split("\n")
            L = srcfile[cand]; text = L[int(ln) - 1].strip()[:70] if int(ln) <= len(L) else ""
    print("%-28s sass %4d  exec %5.2f%%  samples %5.2f%%  %s" % (k, cnt[k], 100 * v / tot, 100 * smp[k] / max(1, ts), text))
