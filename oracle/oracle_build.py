"""TEST INFRASTRUCTURE -- build recipes of the CPU oracle and of the host sources compiled
against it.  Only tests/, __graft_entry__ (build / smoke) and bench.py's CPU-baseline leg use
this; the product package (root_digger_b200/) never imports it.

  oracle/librd_oracle.so             oracle/Makefile
  tests/_build/librd_host_oracle.so  root_digger_b200/host/*.cpp compiled against the oracle
                                     through tests/oracle_shim/rdk.h (-DRD_BACKEND_ORACLE)
"""
from __future__ import annotations

import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from root_digger_b200._build import HOST, _cxx, _newer, _run, build_lbfgsb, host_sources  # noqa: E402

ORACLE = ROOT / "oracle"


def build_oracle(force: bool = False) -> Path:
    """TEST INFRASTRUCTURE: oracle/librd_oracle.so via oracle/Makefile."""
    out = ORACLE / "librd_oracle.so"
    deps = [ORACLE / "rd_oracle.c", ORACLE / "rd_oracle.h", ORACLE / "Makefile"]
    if not force and _newer(out, deps):
        return out
    _run(["make", "-C", ORACLE, "-B" if force else "-s"])
    return out


def build_host_on_oracle(force: bool = False) -> Path:
    """TEST INFRASTRUCTURE: the same host sources compiled against the oracle
    through tests/oracle_shim/rdk.h -> tests/_build/librd_host_oracle.so."""
    oracle = build_oracle()
    build_lbfgsb()
    outdir = ROOT / "tests" / "_build"
    outdir.mkdir(exist_ok=True)
    out = outdir / "librd_host_oracle.so"
    srcs = [s for s in host_sources() if s.exists()]
    shim = ROOT / "tests" / "oracle_shim"
    deps = srcs + list(HOST.glob("*.hpp")) + [shim / "rdk.h", oracle]
    if not force and _newer(out, deps):
        return out
    _run([_cxx(), "-std=c++17", "-O2", "-fPIC", "-fopenmp", "-Wall", "-ffp-contract=off", "-DRD_BACKEND_ORACLE",
          "-I", shim, "-I", ORACLE, "-shared", "-o", out, *srcs, "-L", ORACLE, "-lrd_oracle", "-ldl",
          "-Wl,-rpath," + str(ORACLE), "-Wl,-Bsymbolic"])
    return out


def build_all():
    build_oracle()
    build_host_on_oracle()
