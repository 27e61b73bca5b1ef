"""Partition files and model strings: the reference's known answers
(test/src/msa.cpp:40-283) restated against host/partition_file.cpp, and the ingest path
alignment file + partition file -> model_t (src/main.cpp:513-560) on the oracle backend."""
import math

import numpy as np
import pytest

import fixtures
import oracle_capi
import oracle_build
from root_digger_b200 import capi

RANGES = [(123, 4123), (5122, 12411)]
TAIL = ",PART_0=123-4123, 5122-12411"


@pytest.fixture(scope="module")
def lib():
    oracle_capi.load_oracle().rdo_set_default_mode(oracle_capi.MODE_ENGINE)
    return capi.load_tree_lib(oracle_build.build_host_on_oracle())


def one(lib, line):
    (p,) = capi.parse_partitions(line, lib)
    return p


def test_plain_partition_line(lib):
    """test/src/msa.cpp:40-52"""
    p = one(lib, "DNA, PART_0 = 123-4123")
    assert p["model_name"] == "DNA" and p["partition_name"] == "PART_0" and p["parts"] == [(123, 4123)]
    p = one(lib, "DNA, PART_0 = 123-4123, 5122-12411")
    assert p["parts"] == RANGES
    p = one(lib, "   DNA ,PART_0=123 - 4123 ,5122- 12411  ")
    assert p["parts"] == RANGES and p["subst"] == "DNA"


@pytest.mark.parametrize("model,freq", [("DNA+F", "emperical"), ("DNA+FC", "emperical"), ("DNA+FO", "estimate"),
                                        ("DNA+FE", "equal"), ("DNA+FU{0.25/0.25/0.25/0.25}", "user")])
def test_frequency_options(lib, model, freq):
    """test/src/msa.cpp:53-94"""
    p = one(lib, model + TAIL)
    assert p["model_name"] == model and p["partition_name"] == "PART_0" and p["parts"] == RANGES
    assert p["freq"] == freq


@pytest.mark.parametrize("model,invar,prop", [("DNA+I", "estimate", 0.0), ("DNA+IO", "estimate", 0.0),
                                              ("DNA+IC", "emperical", 0.0), ("DNA+IU{0.25}", "user", 0.25)])
def test_invariant_site_options(lib, model, invar, prop):
    """test/src/msa.cpp:96-131"""
    p = one(lib, model + TAIL)
    assert p["model_name"] == model and p["parts"] == RANGES
    assert p["invar_present"] and p["invar"] == invar and p["invar_prop"] == prop


@pytest.mark.parametrize("model,kind,cat,cats,alpha", [
    ("DNA+G", "estimate", "mean", 4, None), ("DNA+G2", "estimate", "mean", 2, None),
    ("DNA+G2{0.25}", "user", "mean", 2, 0.25), ("DNA+GA", "estimate", "median", 4, None),
    ("DNA+R4", "estimate", "free", 4, None), ("DNA+R2{0.2/0.2}{0.1/0.1}", "estimate", "free", 2, None),
    ("DNA+G16{1.5e+0}", "user", "mean", 16, 1.5)])
def test_rate_heterogeneity_options(lib, model, kind, cat, cats, alpha):
    """test/src/msa.cpp:133-212"""
    p = one(lib, model + TAIL)
    assert p["model_name"] == model and p["parts"] == RANGES
    assert (p["ratehet"], p["cat_type"], p["rate_cats"]) == (kind, cat, cats)
    assert p["alpha_init"] == (alpha is not None)
    if alpha is not None:
        assert p["alpha"] == alpha


def test_all_options_together(lib):
    """test/src/msa.cpp:213-227"""
    p = one(lib, "DNA+G2{0.25}+F+I" + TAIL)
    assert p["model_name"] == "DNA+G2{0.25}+F+I" and p["parts"] == RANGES
    assert (p["ratehet"], p["rate_cats"], p["alpha"]) == ("user", 2, 0.25)
    assert p["invar"] == "estimate" and p["invar_present"] and p["freq"] == "emperical"
    p = one(lib, "UNREST+FO+G4+ASC_LEWIS+M{x}, p = 1-10")
    assert (p["subst"], p["freq"], p["rate_cats"], p["asc"]) == ("UNREST", "estimate", 4, "lewis")
    p = one(lib, "DNA+ASC_STAM{1/2/3/4}+ASC_FELS{7}, p = 1-10")
    assert p["asc"] == "fels"


@pytest.mark.parametrize("line", [
    "DNA PART_0 = 123-4123",        # missing comma        test/src/msa.cpp:229-232
    "DNA, PART_0  123-4123",        # missing =            :233-236
    "DNA, PART_0 = 1234123",        # missing -            :237-240
    "DNA, PART_0 = 123=4123",       # = instead of -       :241-244
    ", PART_0 = 123-4123",          # missing model name   :245-248
    "DNA, PART_0 = 500-100",        # end before begin     src/msa.cpp:466-470
    "DNA+Q, PART_0 = 1-2",          # unknown option
    "DNA+G{, PART_0 = 1-2",         # malformed number
    "DNA+IU{0.1/0.2}, P = 1-2",     # one value expected
    "DNA, PART_0 = 1-2 junk",       # trailing garbage
    "DNA, PART_0 = 1-2, 5",         # a trailing single column needs a range (as in the reference)
])
def test_malformed_lines_raise(lib, line):
    with pytest.raises(ValueError):
        capi.parse_partitions(line, lib)


def test_multi_line_text_and_single_columns(lib):
    ps = capi.parse_partitions("DNA+G, A = 1-100\n\r\n  \nDNA+G2, B = 101-200, 300, 305-310\n", lib)
    assert [p["partition_name"] for p in ps] == ["A", "B"]
    assert ps[1]["parts"] == [(101, 200), (300, 300), (305, 310)]
    assert capi.parse_partitions("", lib) == []


def test_partitioned_alignment_lengths(lib):
    """test/src/msa.cpp:251-283 on 101.phy (compressed when loaded, as msa_t does by default)"""
    path = fixtures.FX / "101.phy"
    L = lambda text: capi.msa_partition_lengths(path, text, True, lib)
    assert L("DNA, PART_0 = 1-100") == [100]
    assert L("DNA, PART_0 = 1-100, 200-300") == [201]
    assert L("DNA, PART_0 = 1-100\nDNA, PART_1 = 200-300") == [100, 101]
    assert L("DNA, PART_0 = 1-100, 500-520\nDNA, PART_1 = 200-300, 400-500") == [121, 202]
    with pytest.raises(ValueError):
        L("DNA, PART_0 = 0-100")          # ranges are 1-based (src/msa.cpp:548-551)
    with pytest.raises(ValueError):
        L("DNA, PART_0 = 1-100000")       # outside the alignment


def test_ingest_alignment_and_partition_file(lib, tmp_path):
    """alignment file + partition file -> multi-partition model_t; a one-partition file that covers
    every column reproduces the unpartitioned model bit for bit"""
    fx = fixtures.load("10.fasta")
    ncol = len(next(iter(fx["alignment"].values())))
    tree = capi.RootedTree(path=str(fx["tree_path"]), lib=lib)
    whole = capi.Model.from_files(tree, fixtures.FX / "10.fasta", None, rate_cats=4, seed=5)
    assert whole.partition_count == 1 and whole.sites() == 991
    pf = tmp_path / "one.part"
    pf.write_text("UNREST+G4, all = 1-%d\n" % ncol)
    one_part = capi.Model.from_files(tree, fixtures.FX / "10.fasta", pf, rate_cats=1, seed=5)
    assert one_part.partition_count == 1 and one_part.sites() == 991
    whole.initialize_partitions(uniform_freqs=True)
    one_part.initialize_partitions(uniform_freqs=True)
    a, b = whole.compute_lh(3), one_part.compute_lh(3)
    assert math.isfinite(a) and a == b
    # two partitions with different numbers of rate categories: logL = sum of the parts
    half = ncol // 2
    pf2 = tmp_path / "two.part"
    pf2.write_text("UNREST+G4, left = 1-%d\nUNREST+G2+FE, right = %d-%d\n" % (half, half + 1, ncol))
    two = capi.Model.from_files(tree, fixtures.FX / "10.fasta", pf2, seed=5)
    assert two.partition_count == 2
    assert two.sites(0) + two.sites(1) >= 991
    two.initialize_partitions(uniform_freqs=True)
    rates, pi = fixtures.FixtureCase.RATES, np.full(4, 0.25)   # the start rates are random per partition
    for q in range(2):
        two.set_params(rates=rates, freqs=pi, part=q)
    lh2 = two.compute_lh(3)
    assert math.isfinite(lh2) and lh2 < 0
    pl = tmp_path / "left.part"
    pl.write_text("UNREST+G4, left = 1-%d\n" % half)
    pr = tmp_path / "right.part"
    pr.write_text("UNREST+G2+FE, right = %d-%d\n" % (half + 1, ncol))
    parts = []
    for f in (pl, pr):
        m = capi.Model.from_files(tree, fixtures.FX / "10.fasta", f, seed=5)
        m.initialize_partitions(uniform_freqs=True)
        m.set_params(rates=rates, freqs=pi)
        parts.append(m.compute_lh(3))
    assert lh2 == parts[0] + parts[1]
    with pytest.raises(RuntimeError):
        capi.Model.from_files(tree, fixtures.FX / "10.fasta", tmp_path / "missing.part")
