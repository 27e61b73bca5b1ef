#!/bin/bash
# usage: sweepcfg.sh variant cfg...
v=$1; shift
for cfg in "$@"; do
  echo "== $v $cfg"
  RDK_ENGINE_LIB=$PWD/root_digger_b200/lib/variants/$v/librdk_b200.so timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --launch-config $cfg 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['logl_root0'], d['best_placement'])
    elif l: print(l[:200])
"
done
