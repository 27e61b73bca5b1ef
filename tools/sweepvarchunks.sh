#!/bin/bash
# usage: sweepvarchunks.sh variant "sites..." "chunks..."
for s in $2; do for c in $3; do
  echo "== $1 sites $s chunks $c"
  RDK_ENGINE_LIB=$PWD/root_digger_b200/lib/variants/$1/librdk_b200.so timeout 120 python bench.py --sites $s --chunks $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['logl_root0'], d['best_placement'], d['full_evaluation']['ms'])
    elif l: print(l[:300])
"
done; done
