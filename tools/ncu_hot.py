"""Per-SASS-instruction executed counts / stall samples of one launch from an .ncu-rep source page.
usage: ncu_hot.py SOURCE.csv [top]   (SOURCE.csv = ncu -i rep --page source --csv ...)
Prints: totals by opcode, and contiguous regions ranked by executed instructions."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[col["Instructions Executed"]].isdigit()]
ex = [int(r[col["Instructions Executed"]]) for r in data]
smp = [int(r[col["# Samples"]]) for r in data]
src = [r[col["Source"]].strip() for r in data]
tot = sum(ex); tots = sum(smp)
print("instructions executed: %.3fG  samples: %d  sass lines: %d" % (tot / 1e9, tots, len(data)))
by = collections.Counter(); bys = collections.Counter()
for s, e, m in zip(src, ex, smp):
    op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
    by[op] += e; bys[op] += m
print("by opcode (executed %, samples %):")
for op, e in by.most_common(28):
    print("  %-10s %6.2f%%  %6.2f%%" % (op, 100.0 * e / tot, 100.0 * bys[op] / max(1, tots)))
# execution-count plateaus: group consecutive lines with similar counts
print("regions (start line, n lines, exec count per line, share of executed, share of samples):")
i = 0; regs = []
while i < len(ex):
    j = i
    while j + 1 < len(ex) and ex[j + 1] > 0 and abs(ex[j + 1] - ex[i]) <= 0.02 * max(ex[i], 1): j += 1
    regs.append((i, j - i + 1, ex[i], sum(ex[i:j + 1]), sum(smp[i:j + 1]))); i = j + 1
regs.sort(key=lambda r: -r[3])
for r in regs[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("  line %5d  n %4d  x %10d  exec %5.2f%%  samples %5.2f%%   %s" % (r[0], r[1], r[2], 100.0 * r[3] / tot, 100.0 * r[4] / max(1, tots), src[r[0]][:60]))
