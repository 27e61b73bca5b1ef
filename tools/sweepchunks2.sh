#!/bin/bash
# usage: sweepchunks2.sh "sites..." "chunks..." "cfg..."   (current engine)
for s in $1; do for c in $2; do for cfg in $3; do
  echo "== sites $s chunks $c cfg $cfg"
  timeout 120 python bench.py --sites $s --chunks $c --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --north-star off --launch-config $cfg 2>&1 | python tools/benchline.py
done; done; done
