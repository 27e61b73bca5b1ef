#!/bin/bash
# usage: sweepchunks.sh "sites..." "chunks..." [extra bench args]
for s in $1; do for c in $2; do
  echo "== sites $s chunks $c $3"
  timeout 120 python bench.py --sites $s --chunks $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['logl_root0'], d['best_placement'], d['full_evaluation']['ms'])
    elif l: print(l[:300])
"
done; done
