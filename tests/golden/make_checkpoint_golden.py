"""Writes tests/golden/checkpoint_v1.ckp: a RootDigger checkpoint laid out by the Python
restatement of the reference's format (tests/test_checkpoint.py, following
src/checkpoint.{hpp,cpp}).  The C++ checkpoint_t must read it and must write the same bytes
for the same inputs (tests/test_checkpoint.py::test_golden_checkpoint_file).

    python tests/golden/make_checkpoint_golden.py
"""
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))

import test_checkpoint as tc  # noqa: E402

OPTIONS = dict(msa="test/data/dna/101.phy", tree="test/data/tree/101.tree", prefix="101.phy",
               model_string="UNREST+G4", rate_cats=(4,), seed=0x5EED0001, min_roots=1, threads=16, exhaustive=True,
               early_stop=2, strategy=2)


def records():
    return [(rid, -21000.0 - 3.25 * rid, 0.125 * (rid % 8), tc.some_params(100 + rid, n=1)) for rid in (0, 7, 42, 198)]


def golden_bytes() -> bytes:
    return tc.expected_header(**OPTIONS) + b"".join(tc.expected_record(*r) for r in records())


if __name__ == "__main__":
    out = HERE / "checkpoint_v1.ckp"
    out.write_bytes(golden_bytes())
    print(out, len(golden_bytes()), "bytes")
