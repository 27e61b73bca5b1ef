#!/bin/bash
# usage: sweepvar.sh "variants..." "sites..." "cfg..." [extra bench args]
for v in $1; do for s in $2; do for cfg in $3; do
  echo "== $v sites $s cfg $cfg"
  RDK_ENGINE_LIB=$PWD/root_digger_b200/lib/variants/$v/librdk_b200.so timeout 120 python bench.py --sites $s --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --launch-config $cfg $4 2>&1 | python tools/benchline.py
done; done; done
