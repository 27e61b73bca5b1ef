"""Instruction mix + stall samples per SASS opcode from an .ncu-rep source page.
usage: ncu_sass_mix.py REPORT.ncu-rep [launch-index]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# split per kernel
blocks, cur = [], []
for line in raw.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = []
    else:
        cur.append(line)
if cur: blocks.append(cur)
rows = list(csv.reader(io.StringIO("\n".join(blocks[which]))))
hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
mix = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[1:]:
    if len(r) < len(hdr): continue
    sass = r[col["Source"]].strip()
    toks = sass.split()
    if not toks: continue
    op = toks[0]
    if op.startswith("@"): op = toks[1] if len(toks) > 1 else op
    op = op.split(".")[0]
    n = int(r[col["Instructions Executed"]] or 0)
    mix[op] += n; tot += n
    samp[op] += int(r[col["# Samples"]] or 0)
print("total warp instrs", tot)
for op, n in mix.most_common(28):
    print("%-10s %12d %5.1f%%  samples %d" % (op, n, 100.0 * n / tot, samp[op]))
