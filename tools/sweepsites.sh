#!/bin/bash
# usage: sweepsites.sh "sites..." "cfg..."   (per-rank shard sizes of a site-sharded cfg2 on one GPU)
for s in $1; do for cfg in $2; do
  echo "== sites $s cfg $cfg"
  timeout 120 python bench.py --sites $s --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --launch-config $cfg 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'])
    elif l: print(l[:200])
"
done; done
