"""SASS <-> CUDA source line map of one kernel of the engine library.
usage: sass_lines.py LIB_OR_OBJ KERNEL_SUBSTRING  -> prints 'index<TAB>line<TAB>sass' for every instruction
(nvdisasm -g line info; the index is the instruction's position in the function, which is also its
row in `ncu --page source --csv`)"""
import re, subprocess, sys, tempfile, os, glob
obj, pat = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
for cubin in sorted(glob.glob(tmp + "/*.cubin")):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    if pat not in out:
        continue
    cur_fn, line, idx, fname = None, 0, 0, ""
    for l in out.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            cur_fn, idx = m.group(1), 0
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = int(m.group(2)); fname = os.path.basename(m.group(1)); continue
        if cur_fn and pat in cur_fn:
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                print("%d\t%s:%d\t%s" % (idx, fname, line, m.group(2).strip())); idx += 1
