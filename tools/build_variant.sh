#!/bin/bash
# usage: tools/build_variant.sh NAME "DEFINES"   -> root_digger_b200/lib/variants/NAME/librdk_b200.so
# (kernel experiments: run bench.py with RDK_ENGINE_LIB pointing at the variant)
set -e
cd "$(dirname "$0")/.."
python - "$1" "$2" <<'PY'
import sys
from root_digger_b200 import _build
out = _build.LIBDIR / "variants" / sys.argv[1] / "librdk_b200.so"
print(_build.build_engine(force=True, out=out, defines=sys.argv[2].split()))
PY
