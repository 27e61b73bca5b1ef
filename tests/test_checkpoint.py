"""checkpoint_t: the reference's "<prefix>.ckp" on-disk format.

The reference's own tests (test/src/checkpoint.cpp) restated, plus a byte-for-byte check
of the file against an independent Python restatement of the format the reference's
templates produce (src/checkpoint.hpp:34-203, src/checkpoint.cpp:11-152): what the C++
writes must be exactly what `expected_*` below build with struct.pack, and a file built
here by hand must be read back by the C++.  (The reference cannot be compiled in this
container -- coraxlib is an absent submodule -- so there is no reference-written file to
pin against; the restatement follows the source line by line.)"""
import os
import struct

import numpy as np
import pytest

import oracle_capi
import oracle_build
from root_digger_b200 import capi

MOD = 65521


@pytest.fixture(scope="module")
def lib():
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    return capi.load_tree_lib(oracle_build.build_host_on_oracle())


# ---- independent restatement of the format -------------------------------------------------
def adler_bytes(data: bytes, a: int, b: int):
    """compute_checksum_components<T> (src/checkpoint.hpp:34-48): `b = b + a % MOD` never
    reduces b, which wraps as a uint32_t"""
    for x in data:
        a = (a + x) % MOD
        b = (b + a % MOD) & 0xFFFFFFFF
    return a, b


def checksum_result(root_id, llh, alpha):
    a, b = adler_bytes(struct.pack("<Qdd", root_id, llh, alpha), 1, 0)
    return ((b << 16) | a) & 0xFFFFFFFF


def checksum_params(params):
    a, b = 1, 0
    for p in params:
        for vec in (p["rates"], p["freqs"], [p["alpha"]], p["weights"]):
            for x in vec:
                a, b = adler_bytes(struct.pack("<d", x), a, b)
        # the variadic overload's last call compute_checksum_components(a, b) resolves to the
        # single-value template: val = a, a0 = b, b0 = 0 (src/checkpoint.hpp:80-85)
        a, b = adler_bytes(struct.pack("<I", a), b, 0)
    return ((b << 16) | a) & 0xFFFFFFFF


def pack_str(s: str) -> bytes:
    return struct.pack("<Q", len(s)) + s.encode()


def expected_header(msa="", tree="", prefix="", model_string="", rate_cats=(1,), seed=0, min_roots=1, threads=0,
                    exhaustive=False, early_stop=0, strategy=2) -> bytes:
    out = pack_str(msa) + pack_str(tree) + pack_str(prefix) + pack_str("") + pack_str("") + pack_str("")
    out += pack_str("") + pack_str("") + pack_str(model_string)
    out += struct.pack("<Q", len(rate_cats))
    for rc in rate_cats:  # ratehet_opts_t(rc): {estimate=1, MEAN=1, rc, alpha_init=false, alpha=1.0}
        out += struct.pack("<iiQB7xd", 1, 1, rc, 0, 1.0)
    out += struct.pack("<QQQ", seed, min_roots, threads)
    out += struct.pack("<ddddd", 0.01, 1e-7, 1e4, 1e-12, 1e-7)
    out += struct.pack("<BBBB", 0, int(exhaustive), 0, 0)
    out += struct.pack("<ii", early_stop, strategy)
    return out + struct.pack("<I", 1)  # CHECKPOINT_WRITE_SUCCESS_FLAG


def pack_doubles(v) -> bytes:
    return struct.pack("<Q", len(v)) + b"".join(struct.pack("<d", x) for x in v)


def expected_record(root_id, llh, alpha, params) -> bytes:
    out = struct.pack("<Qdd", root_id, llh, alpha) + struct.pack("<I", checksum_result(root_id, llh, alpha))
    out += struct.pack("<Q", len(params))
    for p in params:
        out += pack_doubles(p["rates"]) + pack_doubles(p["freqs"]) + pack_doubles([p["alpha"]]) + pack_doubles(p["weights"])
    return out + struct.pack("<I", checksum_params(params))


def some_params(seed, n=2, K=4):
    rng = np.random.default_rng(seed)
    return [{"rates": rng.uniform(1e-4, 1, 12).tolist(), "freqs": rng.dirichlet(np.ones(4)).tolist(),
             "alpha": float(rng.uniform(0.2, 5)), "weights": [1.0 / K] * K} for _ in range(n)]


def new_prefix(tmp_path, name="run"):
    return str(tmp_path / name)


# ---- the reference's tests (test/src/checkpoint.cpp) ---------------------------------------
def test_constructor_creates_the_file(lib, tmp_path):
    """test/src/checkpoint.cpp:26-30"""
    c = capi.Checkpoint(new_prefix(tmp_path), lib)
    assert c.filename.endswith(".ckp") and os.access(c.filename, os.F_OK)
    assert not c.existing


def test_multiple_checkpoints(lib, tmp_path):
    """test/src/checkpoint.cpp:32-38"""
    p = new_prefix(tmp_path)
    c1 = capi.Checkpoint(p, lib)
    assert os.access(c1.filename, os.F_OK)
    c2 = capi.Checkpoint(p, lib)
    assert c2.existing


def test_options_round_trip(lib, tmp_path):
    """test/src/checkpoint.cpp:40-78: default, non-default and changed options"""
    p = new_prefix(tmp_path, "a")
    c1 = capi.Checkpoint(p, lib)
    c1.save_options()
    assert capi.Checkpoint(p, lib).load_options()["equal"]

    p = new_prefix(tmp_path, "b")
    c1 = capi.Checkpoint(p, lib)
    opts = dict(msa="red roses really like to smell good", rate_cats=(1, 1, 3))
    c1.save_options(**opts)
    got = capi.Checkpoint(p, lib).load_options(**opts)
    assert got["equal"] and got["msa"] == opts["msa"] and got["n_rate_cats"] == 3
    changed = dict(opts, msa="this is not the original string")
    assert not capi.Checkpoint(p, lib).load_options(**changed)["equal"]


@pytest.mark.parametrize("count", [1, 1000])
def test_writing_and_reading_results(lib, tmp_path, count):
    """test/src/checkpoint.cpp:80-95"""
    c = capi.Checkpoint(new_prefix(tmp_path), lib)
    c.save_options()
    for _ in range(count):
        c.write(0, 0.0, 0.0, [])
    assert len(c.read_results()) == count


@pytest.mark.parametrize("total", [1, 2, 4, 5, 6, 7, 8, 9, 10])
def test_completed_indicies(lib, tmp_path, total):
    """test/src/checkpoint.cpp:97-118"""
    c = capi.Checkpoint(new_prefix(tmp_path), lib)
    c.save_options()
    for i in range(total):
        c.write(i, 0.0, 0.0, [])
    idx = c.completed_indicies()
    assert len(idx) == total and sorted(idx) == list(range(total))


# ---- the format, byte for byte --------------------------------------------------------------
def test_checksums_match_the_restated_adler_variant(lib):
    c = capi.Checkpoint(None, lib)
    assert c.checksum_result(0, 0.0, 0.0) == checksum_result(0, 0.0, 0.0)
    assert c.checksum_params([]) == 1  # a = 1, b = 0 untouched
    for seed in range(5):
        ps = some_params(seed, n=1 + seed % 3)
        assert c.checksum_params(ps) == checksum_params(ps)
        r = (seed * 977, -12345.678 * (seed + 1), 0.1 * seed)
        assert c.checksum_result(*r) == checksum_result(*r)
    # b really is unreduced: a long parameter list pushes it past 65521 and past 2^16 bits
    big = some_params(99, n=40)
    assert c.checksum_params(big) == checksum_params(big)


def test_file_bytes_equal_the_restated_format(lib, tmp_path):
    p = new_prefix(tmp_path)
    c = capi.Checkpoint(p, lib)
    opts = dict(msa="aln.phy", tree="t.nwk", prefix=p, model_string="UNREST+G4", rate_cats=(4, 1), seed=0xDEADBEEF01,
                min_roots=3, threads=7, exhaustive=True, early_stop=1, strategy=1)
    c.save_options(**opts)
    recs = [(5, -1234.5, 0.25, some_params(1)), (0, -1e7, 1.0, []), (996, -9473046.060349455, 0.5, some_params(2, n=8))]
    for r in recs:
        c.write(*r)
    exp = expected_header(**opts) + b"".join(expected_record(*r) for r in recs)
    assert open(c.filename, "rb").read() == exp


def test_reads_a_file_written_by_the_restated_format(lib, tmp_path):
    """the other direction: a log built by hand (as the reference's writer lays it out) is
    read back, options and parameters included -- a run of this engine resumes from it"""
    p = new_prefix(tmp_path)
    recs = [(3, -100.25, 0.125, some_params(7, n=2)), (11, -99.5, 0.75, some_params(8, n=2))]
    with open(p + ".ckp", "wb") as f:
        f.write(expected_header(msa="x.fasta", seed=42, rate_cats=(4,)))
        for r in recs:
            f.write(expected_record(*r))
    c = capi.Checkpoint(p, lib)
    assert c.existing and not c.needs_cleaning()
    got = c.load_options(msa="x.fasta", seed=42, rate_cats=(4,))
    assert got["equal"] and got["seed"] == 42
    res = c.read_results()
    assert [(r[0], r[1], r[2], r[3]) for r in res] == [(r[0], r[1], r[2], 2) for r in recs]
    for i, r in enumerate(recs):
        for part in range(2):
            q = c.read_params(i, part)
            assert q["rates"].tolist() == r[3][part]["rates"] and q["freqs"].tolist() == r[3][part]["freqs"]
            assert q["alpha"] == r[3][part]["alpha"] and q["weights"].tolist() == r[3][part]["weights"]


def test_golden_checkpoint_file(lib, tmp_path, golden_dir):
    """tests/golden/checkpoint_v1.ckp (written by tests/golden/make_checkpoint_golden.py): read back by
    the C++, and reproduced byte for byte when the C++ writes the same options and records"""
    import importlib.util
    import shutil
    spec = importlib.util.spec_from_file_location("make_checkpoint_golden", golden_dir / "make_checkpoint_golden.py")
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    blob = (golden_dir / "checkpoint_v1.ckp").read_bytes()
    assert blob == mk.golden_bytes()                       # the committed file is what the script writes
    p = new_prefix(tmp_path, "golden")
    shutil.copy(golden_dir / "checkpoint_v1.ckp", p + ".ckp")
    c = capi.Checkpoint(p, lib)
    assert c.existing and not c.needs_cleaning()
    assert c.load_options(**mk.OPTIONS)["equal"]
    recs = mk.records()
    assert c.read_results() == [(r[0], r[1], r[2], 1) for r in recs]
    assert c.completed_indicies() == [0, 7, 42, 198]
    for i, r in enumerate(recs):
        q = c.read_params(i, 0)
        assert q["rates"].tolist() == r[3][0]["rates"] and q["alpha"] == r[3][0]["alpha"]
    w = capi.Checkpoint(new_prefix(tmp_path, "rewrite"), lib)
    w.save_options(**mk.OPTIONS)
    for r in recs:
        w.write(*r)
    assert open(w.filename, "rb").read() == blob


@pytest.mark.parametrize("damage", ["truncate", "flip"])
def test_damaged_tail_is_dropped_and_cleaned(lib, tmp_path, damage):
    """src/checkpoint.cpp:318-324 ("resume with what we can"), needs_cleaning :329-362, clean :165-190"""
    p = new_prefix(tmp_path)
    c = capi.Checkpoint(p, lib)
    c.save_options()
    for i in range(4):
        c.write(i, -1.0 * i, 0.5, some_params(i, n=1))
    size = os.path.getsize(c.filename)
    rec = len(expected_record(3, -3.0, 0.5, some_params(3, n=1)))
    with open(c.filename, "r+b") as f:
        if damage == "truncate":
            f.truncate(size - 9)          # a torn last record
        else:
            f.seek(size - rec + 10)       # a flipped bit in the last result
            b = f.read(1)
            f.seek(size - rec + 10)
            f.write(bytes([b[0] ^ 0x40]))
    c2 = capi.Checkpoint(p, lib)
    assert c2.needs_cleaning()
    assert c2.completed_indicies() == [0, 1, 2]
    c2.clean()
    assert not c2.needs_cleaning() and c2.completed_indicies() == [0, 1, 2]
    assert os.path.getsize(c2.filename) == size - rec and not os.path.exists(c2.filename + ".bak")
    c2.write(3, -3.0, 0.5, some_params(3, n=1))   # the log keeps growing after a clean
    assert capi.Checkpoint(p, lib).completed_indicies() == [0, 1, 2, 3]


def test_in_memory_log(lib):
    c = capi.Checkpoint(None, lib)
    c.write(4, -2.0, 0.5, some_params(0, n=1))
    assert c.read_results() == [(4, -2.0, 0.5, 1)] and not c.needs_cleaning()


def test_exhaustive_search_resumes_from_its_checkpoint(lib, tmp_path):
    """a run interrupted after some branches picks up the remaining ones only
    (assign_indicies_by_rank_exhaustive, reference src/model.cpp:1934-1960) and ends with a
    result for every branch, the finished ones taken from the file"""
    from root_digger_b200 import synth
    rng = np.random.default_rng(5)
    tree_text = "((a:0.1,b:0.2):0.05,(c:0.15,d:0.1):0.3,e:0.2);"
    aln = {t: "".join(rng.choice(list("ACGT"), 60)) for t in "abcde"}
    tol = (1e-2, 1e-2, 1e-2, 1e13)

    def model():
        m = capi.Model(capi.RootedTree(tree_text, lib=lib), aln, rate_cats=1, compress=True, invariant_sites=True,
                       seed=12345)
        m.initialize_partitions(uniform_freqs=False)
        return m

    p = new_prefix(tmp_path)
    first = model()
    n = first.root_count
    first.set_checkpoint(p)
    ids1, llh1, alpha1 = first.exhaustive_search(*tol, rank=0, num_tasks=2)   # some branches, then "crash"
    assert 0 < len(ids1) < n
    ckp = capi.Checkpoint(p, lib)
    assert ckp.existing and sorted(ckp.completed_indicies()) == sorted(ids1.tolist())
    size1 = os.path.getsize(ckp.filename)
    again = model()
    again.set_checkpoint(p)
    ids2, llh2, alpha2 = again.exhaustive_search(*tol)
    assert sorted(ids2.tolist()) == list(range(n)) and np.isfinite(llh2).all()
    # the records of the first run are still the first records of the log, untouched
    k = len(ids1)
    assert ids2[:k].tolist() == ids1.tolist() and np.array_equal(llh2[:k], llh1) and np.array_equal(alpha2[:k], alpha1)
    assert os.path.getsize(ckp.filename) > size1 and not ckp.needs_cleaning()
    stored = ckp.read_params(0, 0, K=1)
    assert (stored["rates"] > 0).all() and (stored["freqs"] > 0).all()   # the optimiser's raw variables
    # a third run has nothing left to do
    third = model()
    third.set_checkpoint(p)
    ids3, _, _ = third.exhaustive_search(*tol)
    assert ids3.tolist() == ids2.tolist()


@pytest.mark.parametrize("dummy", [0, 1, 2, 4, 8])
def test_assign_indicies_against_the_checkpoint(lib, tmp_path, dummy):
    """test/src/model.cpp:448-551: 10.fasta has 2n-3 = 17 rootings; with `dummy` finished roots in the
    checkpoint file, the search assignment hands out max(k - dummy, 0) of k requested start roots (or
    throws when the file holds more than were asked for) and the exhaustive assignment the 17 - dummy
    remaining ones, never a finished one -- for every start-root strategy, and split over ranks"""
    import fixtures
    rng = np.random.default_rng(dummy)
    done = rng.permutation(17)[:dummy].tolist()
    p = new_prefix(tmp_path)
    ckp = capi.Checkpoint(p, lib)
    ckp.save_options()
    for rid in done:
        ckp.write(int(rid), 0.0, 0.0, [])
    fx = fixtures.load("10.fasta")
    tree = capi.RootedTree(path=str(fx["tree_path"]), lib=lib)
    m = capi.Model(tree, fx["alignment"], rate_cats=1, compress=True, seed=99)
    m.initialize_partitions(uniform_freqs=True)
    m.set_checkpoint(p)
    assert m.root_count == 17
    for strategy in ("random", "midpoint", "modified_mad"):
        for k in (1, 2, 3, 4, 5):
            if k - dummy >= 0:
                got = m.assign_indicies("search", min_roots=k, root_ratio=0.0, strategy=strategy)
                assert len(got) == k - dummy and not set(got) & set(done) and len(set(got)) == len(got)
            else:
                with pytest.raises(RuntimeError):
                    m.assign_indicies("search", min_roots=k, root_ratio=0.0, strategy=strategy)
    got = m.assign_indicies("exhaustive")
    assert sorted(got) == sorted(set(range(17)) - set(done))
    # the same work split over 3 ranks (src/model.cpp:1899-1907): disjoint, complete, balanced
    parts = [m.assign_indicies("exhaustive", rank=r, num_tasks=3) for r in range(3)]
    assert sorted(sum(parts, [])) == sorted(got) and max(map(len, parts)) - min(map(len, parts)) <= 1
