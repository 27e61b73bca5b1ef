"""Host traversal scheduler against the reference's known-answer tests
(test/src/tree.cpp) -- the golden vectors that exist for this path."""
import numpy as np
import pytest

import fixtures
from root_digger_b200.capi import RootedTree
from root_digger_b200 import synth

SINGLE = "((a:.1,b:.1)n1:.05,(c:.1,d:.1)n2:.5);"


def test_generate_operations_known_tree():
    """test/src/tree.cpp:142-180"""
    t = RootedTree(SINGLE)
    ops, pm, br = t.generate_operations(t.root_id("n2"), 0.5)
    assert [o.astuple() for o in ops] == [(4, 0, 0, 0, -1, 1, 1, -1), (5, 1, 2, 2, -1, 3, 3, -1),
                                          (6, 2, 4, 4, 0, 5, 5, 1)]
    assert t.root_clv_index == 6 and t.root_scaler_index == 2
    assert len(pm) == 6 and sorted(pm.tolist()) == [0, 1, 2, 3, 4, 5]


def test_generate_derivative_operations_known_tree():
    """test/src/tree.cpp:182-212"""
    t = RootedTree(SINGLE)
    op, pm, br = t.generate_derivative_operations(t.root_id("n2"), 0.5)
    assert op.astuple() == (6, 2, 4, 4, 0, 5, 5, 1)
    assert pm.tolist() == [4, 5] and br.tolist() == [0.275, 0.275]


def test_newick_after_rerooting():
    """test/src/tree.cpp:225-292"""
    t = RootedTree(SINGLE)
    assert t.root_count == 5
    exp = {
        ("b", .25): "(b:0.025000,((c:0.100000,d:0.100000)n2:0.550000,a:0.100000)n1:0.075000);",
        ("b", .75): "(b:0.075000,((c:0.100000,d:0.100000)n2:0.550000,a:0.100000)n1:0.025000);",
        ("a", .25): "(a:0.025000,(b:0.100000,(c:0.100000,d:0.100000)n2:0.550000)n1:0.075000);",
        ("a", .75): "(a:0.075000,(b:0.100000,(c:0.100000,d:0.100000)n2:0.550000)n1:0.025000);",
        ("n2", .25): "((c:0.100000,d:0.100000)n2:0.137500,(a:0.100000,b:0.100000)n1:0.412500);",
        ("n2", .75): "((c:0.100000,d:0.100000)n2:0.412500,(a:0.100000,b:0.100000)n1:0.137500);",
        ("c", .25): "(c:0.025000,(d:0.100000,(a:0.100000,b:0.100000)n1:0.550000)n2:0.075000);",
        ("c", .75): "(c:0.075000,(d:0.100000,(a:0.100000,b:0.100000)n1:0.550000)n2:0.025000);",
        ("d", .25): "(d:0.025000,((a:0.100000,b:0.100000)n1:0.550000,c:0.100000)n2:0.075000);",
        ("d", .75): "(d:0.075000,((a:0.100000,b:0.100000)n1:0.550000,c:0.100000)n2:0.025000);",
    }
    for (label, ratio), s in exp.items():
        t.root_by(t.root_id(label), ratio)
        assert t.newick() == s


def test_derivative_op_is_last_full_op():
    """test/src/tree.cpp:298-334"""
    t = RootedTree(SINGLE)
    for rid in range(t.root_count):
        ops, _, _ = t.generate_operations(rid, 0.5)
        op, _, _ = t.generate_derivative_operations(rid, 0.5)
        assert op.astuple() == ops[-1].astuple()


def test_root_update_operations():
    """test/src/tree.cpp:410-433"""
    t = RootedTree(SINGLE)
    t.root_by(t.root_id("a"))
    ops, pm, br = t.generate_root_update_operations(t.root_id("d"))
    assert (len(ops), len(pm), len(br)) == (3, 4, 4)
    t2 = RootedTree(SINGLE)
    t2.root_by(t2.root_id("b"))
    ops, pm, br = t2.generate_root_update_operations(t2.root_id("b"))
    assert (len(ops), len(pm), len(br)) == (0, 0, 0)


def test_annotations_and_midpoint_on_10_tree():
    """test/src/tree.cpp:347-364, 435-443"""
    t = RootedTree(path=fixtures.FX / "10.tree")
    assert t.root_count == 17
    for rid in range(t.root_count):
        t.annotate_branch(rid, "foo", "bar")
        t.annotate_branch(rid, "fizz", "buzz")
    a = "[&&NHX:foo=bar:fizz=buzz]"
    assert t.newick() == (
        f"(((j:0.854700{a},((h:0.983500{a},a:0.224900{a}):0.416200{a},(c:0.540900{a},f:0.422200{a}):0.785300{a})"
        f":0.614100{a}):0.446100{a},g:0.487400{a}):0.825200{a},((i:0.569700{a},e:0.366600{a}):0.602800{a},"
        f"b:0.445900{a}):0.099300{a},d:0.639600{a});")
    t2 = RootedTree(path=fixtures.FX / "10.tree")
    t2.root_by(t2.rank_roots("midpoint")[0])
    assert t2.newick(False) == (
        "((j:0.854700,((h:0.983500,a:0.224900):0.416200,(c:0.540900,f:0.422200):0.785300):0.614100):0.223050,"
        "(g:0.487400,(((i:0.569700,e:0.366600):0.602800,b:0.445900):0.099300,d:0.639600):0.825200):0.223050);")


def test_annotations_survive_rerooting():
    """test/src/tree.cpp:385-409 ("annotations, all roots") and :366-383 (the hidden "moving root"
    case): annotate every branch -- while re-rooting on each, or from one fixed root -- then root on a
    tip, unroot and print: the annotations stay with their branches"""
    a = "[&&NHX:foo=bar:fizz=buzz]"
    t = RootedTree(path=fixtures.FX / "10.tree")
    for rid in range(t.root_count):
        t.root_by(rid)
        t.annotate_branch(rid, "foo", "bar")
        t.annotate_branch(rid, "fizz", "buzz")
    t.root_by(t.root_id("a"))
    t.unroot()
    assert t.newick() == (
        f"(a:0.224900{a},((c:0.540900{a},f:0.422200{a}):0.785300{a},((g:0.487400{a},(((i:0.569700{a},"
        f"e:0.366600{a}):0.602800{a},b:0.445900{a}):0.099300{a},d:0.639600{a}):0.825200{a}):0.446100{a},"
        f"j:0.854700{a}):0.614100{a}):0.416200{a},h:0.983500{a});")
    t = RootedTree(path=fixtures.FX / "10.tree")
    t.root_by(0)
    for rid in range(t.root_count):
        t.annotate_branch(rid, "foo", "bar")
        t.annotate_branch(rid, "fizz", "buzz")
    t.root_by(t.root_id("b"))
    t.unroot()
    assert t.newick() == (
        f"(b:0.445900{a},(d:0.639600{a},((j:0.854700{a},((h:0.983500{a},a:0.224900{a}):0.416200{a},"
        f"(c:0.540900{a},f:0.422200{a}):0.785300{a}):0.614100{a}):0.446100{a},g:0.487400{a}):0.825200{a})"
        f":0.099300{a},(i:0.569700{a},e:0.366600{a}):0.602800{a});")


@pytest.mark.parametrize("name", ["single", "10", "101"])
def test_construction_copy_labels_and_operations_on_the_bundled_trees(name):
    """test/src/tree.cpp:18-140, :214-223: every root location has a positive saved branch length, two
    constructions agree root by root, a copy has the same shape, the label map covers the tips, every
    root yields a non-empty schedule with positive branch lengths, and root_by / unroot cycles through
    every root"""
    path = fixtures.FX / (name + ".tree")
    t1, t2 = RootedTree(path=path), RootedTree(path=path)
    n = t1.tip_count
    assert t1.root_count == 2 * n - 3 > 0 and t1.newick() == t2.newick()
    for rid in range(t1.root_count):
        a, b = t1.root_info(rid), t2.root_info(rid)
        assert a[0] > 0.0 and a == b
    c = t1.copy()
    assert (c.tip_count, c.inner_count, c.branch_count, c.root_count) == (
        t1.tip_count, t1.inner_count, t1.branch_count, t1.root_count)
    assert sorted(t1.tip_index(t1.tip_label(i)) for i in range(n)) == list(range(n))
    for rid in range(t1.root_count):
        ops, pm, br = t1.generate_operations(rid, 0.5)
        assert len(ops) == n - 1 and len(pm) == len(br) == 2 * n - 2 and (br > 0.0).all()
    for rid in range(t1.root_count):
        t1.root_by(rid)
        assert t1.rooted
        t1.unroot()
        assert not t1.rooted
    assert c.newick() == t2.newick()
    with pytest.raises(Exception):
        RootedTree(path="not_a_tree_file")


def test_sanity_check_trees():
    """test/src/tree.cpp:336-345"""
    assert not RootedTree(path=fixtures.FX / "sanity_check1.tree").sanity_check()
    assert not RootedTree(path=fixtures.FX / "sanity_check2.tree").sanity_check()
    assert RootedTree(path=fixtures.FX / "sanity_check3.tree").sanity_check()


def test_bad_input_is_rejected():
    with pytest.raises(ValueError):
        RootedTree(path="not_a_tree_file")
    with pytest.raises(ValueError):
        RootedTree("(a:1,b:1);")
    with pytest.raises(ValueError):
        RootedTree("((a:1,b:1,c:1,d:1):1,e:1,f:1);")


@pytest.mark.parametrize("n", [3, 4, 17, 101, 500])
def test_schedule_invariants_on_random_trees(n):
    top = synth.random_tree(n, n)
    t = RootedTree(synth.to_newick(top))
    assert t.tip_count == n and t.root_count == 2 * n - 3 and t.branch_count == 2 * n - 2
    for rid in range(0, t.root_count, max(1, t.root_count // 7)):
        for ratio in (0.0, 0.3, 1.0):
            ops, pm, br = t.generate_operations(rid, ratio)
            assert len(ops) == n - 1 and len(pm) == 2 * n - 2
            assert sorted(pm.tolist()) == list(range(2 * n - 2))       # every branch exactly once
            assert (br >= 0).all()
            written = set(range(n))
            for o in ops:                                               # children before parents
                assert o.child1_clv_index in written and o.child2_clv_index in written
                assert o.parent_clv_index >= n and o.parent_clv_index not in written
                written.add(o.parent_clv_index)
                for c, s in ((o.child1_clv_index, o.child1_scaler_index), (o.child2_clv_index, o.child2_scaler_index)):
                    assert (s == -1) == (c < n)
            assert ops[-1].parent_clv_index == t.root_clv_index == 2 * n - 2
            assert ops[-1].parent_scaler_index == t.root_scaler_index == n - 2
            saved, _, _ = t.root_info(rid)
            assert abs(br[pm.tolist().index(ops[-1].child1_matrix_index)] - saved * ratio) < 1e-15
    # a copy of an unrooted tree keeps the root ids
    t.unroot()
    c = t.copy()
    assert [c.root_info(i)[0] for i in range(c.root_count)] == [t.root_info(i)[0] for i in range(t.root_count)]


@pytest.mark.parametrize("chunks,begin,end", [(3, None, None), (16, None, None), (4, 10, 61), (5, 0, 7)])
def test_chunked_sweep_schedule_on_the_oracle(chunks, begin, end):
    """RootedTree.generate_chunked_sweep_operations (what bench.py feeds
    rdk_sweep_root_placements_chunks): chunks use disjoint spare buffers, cover the requested
    root positions once, and -- run in order on the oracle -- give the bits of the unchunked sweep"""
    import numpy as np
    from cases import Case, compute_lh, same_bits
    from oracle_capi import MODE_ENGINE, OraclePartition
    case = Case(40, 300, 4, seed=31, data="ambiguous", weights="random")
    lay = case.tree.sweep_layout(chunks)
    o = OraclePartition(case.n, case.S, case.K, clv_buffers=lay["clv_buffers"], scale_buffers=lay["scale_buffers"],
                        prob_matrices=lay["prob_matrices"])
    case.setup(o)
    compute_lh(o, case.full_schedule(5, 0.4), case.root_clv, case.root_scaler)
    *sw, pos = case.tree.generate_sweep_operations(begin, end, layout=lay)
    want = np.empty(case.tree.root_count)
    want[pos] = o.sweep_root_placements(*sw, case.root_clv, case.root_scaler, mode=MODE_ENGINE)
    compute_lh(o, case.full_schedule(5, 0.4), case.root_clv, case.root_scaler)
    *csw, cpos, coff = case.tree.generate_chunked_sweep_operations(begin, end, layout=lay)
    b, e = (0 if begin is None else begin), (case.tree.root_count if end is None else end)
    assert sorted(cpos.tolist()) == list(range(b, e)) and coff[0] == 0 and coff[-1] == len(cpos)
    assert len(coff) - 1 == max(1, min(chunks, (e - b) // 8))
    # spare buffers written by different chunks are disjoint
    op_off, ops = csw[3], csw[4]
    written = []
    for c in range(len(coff) - 1):
        w = {ops[i].parent_clv_index for i in range(op_off[coff[c]], op_off[coff[c + 1]])} - {case.root_clv}
        assert all(lay["clv0"] + c * lay["extra"] <= x < lay["clv0"] + (c + 1) * lay["extra"] for x in w)
        written.append(w)
    assert all(written[i].isdisjoint(written[j]) for i in range(len(written)) for j in range(i))
    got = np.empty(case.tree.root_count)
    got[cpos] = o.sweep_root_placements(*csw, case.root_clv, case.root_scaler, mode=MODE_ENGINE)
    assert same_bits(got[b:e], want[b:e])
