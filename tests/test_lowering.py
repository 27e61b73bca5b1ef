"""Data flow of the lowering (root_digger_b200/csrc/rdk_lower.hpp), checked without a GPU.

The engine turns recorded CLV operations into the instructions the program kernel walks: it
orders the two children canonically, keeps ONE CLV value in the warps' registers (`v`), inserts
register loads where `v` does not hold the child an operation needs, and drops the stores nobody
reads back.  Here the lowered program is interpreted SYMBOLICALLY -- a CLV value is a nested
tuple naming how it was computed -- next to the straightforward execution of the recorded
operations (reference semantics: corax_update_clvs executes its operations in array order,
src/model.cpp:402,440,461,851), and every evaluated value and every buffer whose content is
required afterwards must be the same term.  The product of the two children's terms is
commutative (bit for bit: one rounded multiplication per state), so a term is normalised by
sorting its two factors.
"""
import ctypes as C
import random

import numpy as np
import pytest

from root_digger_b200 import capi

R_WRITE, R_EVAL, R_LOADONLY = 1, 2, 4
F = dict(Tip1=1, Tip2=2, Write=4, WriteS=8, Eval=16, Scale=32, LoadV=64, Nop=128, EvalV=256, Cnt1=512,
         Cnt2V=1024, EvalScaler=2048, LoadV2=4096, Cnt2M=8192)
NONE = 0xFFFFFFFF


def lower(tips, ops, chunk_off=None, discard=False, scratch=None):
    L = capi.load_engine()
    L.rdk_debug_lower_program.restype = C.c_int
    n = len(ops)
    arr = np.ascontiguousarray(np.array(ops, dtype=np.int64).astype(np.int32).reshape(n, 10))
    out = np.zeros((4 * n + 16, 10), dtype=np.int32)
    nch = len(chunk_off) - 1 if chunk_off else 1
    co = np.ascontiguousarray(np.array(chunk_off if chunk_off else [0, n], dtype=np.uint32))
    oco = np.zeros(nch + 1, dtype=np.uint32)
    sc = np.ascontiguousarray(np.array(scratch, dtype=np.uint8)) if scratch is not None else None
    cnt = L.rdk_debug_lower_program(
        C.c_uint(tips), C.c_uint(n), arr.ctypes.data_as(C.POINTER(C.c_int)), C.c_uint(nch),
        co.ctypes.data_as(C.POINTER(C.c_uint)), C.c_int(1 if discard else 0),
        sc.ctypes.data_as(C.POINTER(C.c_ubyte)) if sc is not None else None, C.c_uint(0 if sc is None else len(sc)),
        out.ctypes.data_as(C.POINTER(C.c_int)), C.c_uint(out.shape[0]), oco.ctypes.data_as(C.POINTER(C.c_uint)))
    assert cnt >= 0
    return [tuple(int(x) & 0xFFFFFFFF if j in (1, 3, 5) else int(x) for j, x in enumerate(row)) for row in out[:cnt]], \
        [int(x) for x in oco]


def term(a, b):
    return ("mul",) + tuple(sorted((a, b), key=repr))


class Mem:
    """CLV buffers and scale buffers holding symbolic values"""

    def __init__(self, tips, n_clv, n_sc):
        self.tips = tips
        self.clv = {i: ("init", i) for i in range(tips, n_clv)}
        self.sc = {i: ("sinit", i) for i in range(n_sc)}


def run_reference(mem, ops):
    """recorded operations executed eagerly, in order"""
    evals = {}
    for (parent, pscale, c1, c2, s1, s2, pm1, pm2, flags, slot) in ops:
        if flags & R_LOADONLY:
            evals[slot] = ("eval", mem.clv[c1], mem.sc[s1] if s1 >= 0 else 0)
            continue

        def child(c, pm):
            return ("tip", c, pm) if c < mem.tips else ("mv", pm, mem.clv[c])
        r = term(child(c1, pm1), child(c2, pm2))
        cnt = ("cnt", mem.sc[s1] if s1 >= 0 else 0, mem.sc[s2] if s2 >= 0 else 0, r) if pscale >= 0 else 0
        cnt = normalise_cnt(cnt)
        if pscale >= 0:
            r = ("scaled", r)
        if flags & R_WRITE:
            mem.clv[parent] = r
            if pscale >= 0:
                mem.sc[pscale] = cnt
        if flags & R_EVAL:
            evals[slot] = ("eval", r, cnt if pscale >= 0 else 0)
    return evals


def normalise_cnt(c):
    if c == 0:
        return 0
    return ("cnt",) + tuple(sorted(c[1:3], key=repr)) + (c[3],)


def run_lowered(mem, prog):
    """the kernel's semantics: registers v / vcnt, operands loaded from memory"""
    evals = {}
    v, vcnt = ("garbage",), ("garbage",)
    prev_written = (None, None)
    for (fl, parent, pscale, c1, s1, c2, pm1, pm2, slot, s2) in prog:
        loads_c1 = not (fl & (F["Tip1"] | F["Nop"]))
        # operands are fetched while the previous instruction computes: they must not be what it stores
        if loads_c1:
            assert c1 != prev_written[0], "operand prefetch would race with the previous store"
        if fl & F["Cnt1"]:
            assert s1 != prev_written[1], "scaler prefetch would race with the previous store"
        cnt1 = mem.sc[s1] if fl & F["Cnt1"] else 0
        prev_written = (None, None)
        if fl & F["LoadV2"]:  # issued after the previous instruction's stores
            v, vcnt = mem.clv[c2], (mem.sc[s2] if fl & F["Cnt2M"] else 0)
        if fl & F["LoadV"]:
            v, vcnt = mem.clv[c1], cnt1
        elif not (fl & F["Nop"]):
            a = ("tip", c1, pm1) if fl & F["Tip1"] else ("mv", pm1, mem.clv[c1])
            b = ("tip", c2, pm2) if fl & F["Tip2"] else ("mv", pm2, v)
            r = term(a, b)
            cnt2 = vcnt if fl & F["Cnt2V"] else 0
            cnt = 0
            if fl & F["Scale"]:
                cnt = normalise_cnt(("cnt", cnt1, cnt2, r))
                r = ("scaled", r)
            if fl & F["Eval"]:
                assert not (fl & (F["Write"] | F["WriteS"]))
                evals[slot] = ("eval", r, cnt if fl & F["EvalScaler"] else 0)
            else:
                if fl & F["Write"]:
                    mem.clv[parent] = r
                if fl & F["WriteS"]:
                    mem.sc[pscale] = cnt
                prev_written = (parent if fl & F["Write"] else None, pscale if fl & F["WriteS"] else None)
                v, vcnt = r, cnt
        if fl & F["EvalV"]:
            evals[slot] = ("eval", v, vcnt if fl & F["EvalScaler"] else 0)
    return evals


def random_tree_ops(rng, tips, with_eval=True):
    """a post-order traversal of a random rooted binary tree, corax-style indices"""
    nodes = list(range(tips))
    next_clv, next_sc, next_pm = tips, 0, 0
    scaler_of = {}
    ops = []
    rng.shuffle(nodes)
    # random sequential joining: not a post-order of one tree in general, so forwarding, register
    # loads and plain loads all occur
    while len(nodes) > 1:
        i, j = rng.sample(range(len(nodes)), 2)
        a, b = nodes[i], nodes[j]
        for x in sorted((i, j), reverse=True):
            nodes.pop(x)
        p, ps = next_clv, next_sc
        next_clv += 1
        next_sc += 1
        ops.append((p, ps, a, b, scaler_of.get(a, -1), scaler_of.get(b, -1), next_pm, next_pm + 1, R_WRITE, 0))
        next_pm += 2
        scaler_of[p] = ps
        nodes.append(p)
    if with_eval:
        last = list(ops[-1])
        last[8] |= R_EVAL
        ops[-1] = tuple(last)
    return ops, next_clv, next_sc


@pytest.mark.parametrize("seed", range(12))
def test_random_traversals_keep_their_data_flow(seed):
    rng = random.Random(seed)
    tips = rng.choice([2, 3, 5, 9, 17, 40])
    ops, n_clv, n_sc = random_tree_ops(rng, tips)
    prog, _ = lower(tips, ops)
    ref, low = Mem(tips, n_clv, n_sc), Mem(tips, n_clv, n_sc)
    assert run_reference(ref, ops) == run_lowered(low, prog)
    assert ref.clv == low.clv and ref.sc == low.sc  # nothing is scratch: every buffer must be stored


def test_post_order_forwards_and_discard_drops_every_store():
    rng = random.Random(5)
    tips = 33
    # a caterpillar: every inner child is produced by the instruction just before its parent
    ops, clv, sc, pm = [], tips, 0, 0
    prev, prev_sc = 0, -1
    for t in range(1, tips):
        ops.append((clv, sc, prev, t, prev_sc, -1, pm, pm + 1, R_WRITE, 0))
        prev, prev_sc = clv, sc
        clv, sc, pm = clv + 1, sc + 1, pm + 2
    ops[-1] = ops[-1][:8] + (R_WRITE | R_EVAL, 0)
    prog, _ = lower(tips, ops)
    assert len(prog) == len(ops)  # no register load needed
    assert all(not (fl & (F["LoadV"] | F["LoadV2"])) for fl, *_ in prog)
    ref, low = Mem(tips, clv, sc), Mem(tips, clv, sc)
    assert run_reference(ref, ops) == run_lowered(low, prog)
    assert ref.clv == low.clv
    # the same traversal as a lazily materialised evaluation: no store survives, same value
    prog2, _ = lower(tips, ops, discard=True)
    assert all(not (fl & (F["Write"] | F["WriteS"])) for fl, *_ in prog2)
    low2 = Mem(tips, clv, sc)
    assert run_lowered(low2, prog2) == run_reference(Mem(tips, clv, sc), ops)
    del rng


def directed_sweep_ops(rng, tips):
    """ops of a directed-CLV sweep (host/tree.cpp generate_sweep_operations) over a random tree,
    spare buffers per depth; returns (setup ops, sweep ops, n_clv, n_sc, spare clv range)"""
    # build a random rooted binary tree
    class N:
        pass
    leaves = []
    for t in range(tips):
        n = N()
        n.clv, n.sc, n.kids, n.pm = t, -1, None, t
        leaves.append(n)
    pool = leaves[:]
    clv, sc = tips, 0
    post = []
    while len(pool) > 1:
        i, j = rng.sample(range(len(pool)), 2)
        a, b = pool[i], pool[j]
        for x in sorted((i, j), reverse=True):
            pool.pop(x)
        n = N()
        n.clv, n.sc, n.kids, n.pm = clv, sc, (a, b), clv
        clv, sc = clv + 1, sc + 1
        pool.append(n)
    root = pool[0]

    def walk(n):
        if n.kids:
            walk(n.kids[0])
            walk(n.kids[1])
            post.append((n.clv, n.sc, n.kids[0].clv, n.kids[1].clv, n.kids[0].sc, n.kids[1].sc,
                         n.kids[0].pm, n.kids[1].pm, R_WRITE, 0))
    walk(root)
    setup = post
    spare0, spare_sc0 = clv, sc
    root_clv, root_sc = clv + 64, sc + 64
    ops = []
    slot = [0]
    pmx = [10_000]

    def placement(below, above):
        ops.append((root_clv, root_sc, below[0], above[0], below[1], above[1], pmx[0], pmx[0] + 1, R_EVAL, slot[0]))
        pmx[0] += 2
        slot[0] += 1

    def descend(n, up, edge_pm, depth):
        if not n.kids:
            return
        for i in (0, 1):
            child, sib = n.kids[i], n.kids[1 - i]
            U = (spare0 + depth, spare_sc0 + depth)
            ops.append((U[0], U[1], up[0], sib.clv, up[1], sib.sc, edge_pm, sib.pm, R_WRITE, 0))
            placement((child.clv, child.sc), U)
            descend(child, U, child.pm, depth + 1)

    l, r = root.kids
    placement((l.clv, l.sc), (r.clv, r.sc))
    descend(l, (r.clv, r.sc), 9_999, 0)
    descend(r, (l.clv, l.sc), 9_999, 0)
    return setup, ops, root_clv + 1, root_sc + 1, slot[0]


@pytest.mark.parametrize("seed", range(8))
def test_directed_sweep_with_scratch_buffers(seed):
    rng = random.Random(100 + seed)
    tips = rng.choice([4, 7, 12, 30])
    setup, ops, n_clv, n_sc, n_eval = directed_sweep_ops(rng, tips)
    ref = Mem(tips, n_clv, n_sc)
    run_reference(ref, setup)
    low = Mem(tips, n_clv, n_sc)
    low.clv, low.sc = dict(ref.clv), dict(ref.sc)
    before = dict(ref.clv)
    want = run_reference(ref, ops)
    prog, _ = lower(tips, ops, discard=True)
    got = run_lowered(low, prog)
    assert len(want) == n_eval and got == want
    # the partition's own CLVs are untouched
    for c in range(tips, tips + len(setup)):
        assert low.clv[c] == before[c]
    # the evaluation leaves v alone: the first child below an inner edge finds its directed CLV in
    # registers, and the directed CLVs towards tips are never stored
    n_ops = sum(1 for o in ops if o[8] & R_WRITE)
    stored = sum(1 for fl, *_ in prog if fl & F["Write"])
    loadv = sum(1 for fl, *_ in prog if fl & (F["LoadV"] | F["LoadV2"]))
    assert stored < n_ops and (loadv < n_ops or tips < 12)
    # in chunks: every chunk starts with unknown registers
    half = len(ops) // 2
    while half < len(ops) and not (ops[half - 1][8] & R_EVAL):
        half += 1
    if 0 < half < len(ops):
        progc, coff = lower(tips, ops, chunk_off=[0, half, len(ops)])
        low2 = Mem(tips, n_clv, n_sc)
        ref2 = Mem(tips, n_clv, n_sc)
        run_reference(ref2, setup)
        low2.clv, low2.sc = dict(ref2.clv), dict(ref2.sc)
        got2 = {}
        for c in range(2):
            got2.update(run_lowered(low2, progc[coff[c]:coff[c + 1]]))
        assert got2 == want


def test_degenerate_operand_patterns():
    tips = 4
    # both children the same inner CLV, straight after it was produced; and a stored-CLV evaluation
    ops = [
        (4, 0, 0, 1, -1, -1, 0, 1, R_WRITE, 0),
        (5, 1, 4, 4, 0, 0, 2, 3, R_WRITE, 0),
        (0xFFFFFFFF, -1, 5, 0xFFFFFFFF, 1, -1, 0, 0, R_LOADONLY | R_EVAL, 0),
        (6, 2, 5, 4, 1, 0, 4, 5, R_WRITE | R_EVAL, 1),
        (0xFFFFFFFF, -1, 4, 0xFFFFFFFF, 0, -1, 0, 0, R_LOADONLY | R_EVAL, 2),
    ]
    ops = [tuple(-1 if x == 0xFFFFFFFF else x for x in o) for o in ops]
    prog, _ = lower(tips, ops)
    ref, low = Mem(tips, 8, 4), Mem(tips, 8, 4)
    want = run_reference(ref, [tuple(o) for o in ops])
    got = run_lowered(low, prog)
    assert got == want and ref.clv == low.clv and ref.sc == low.sc
    assert any(fl & F["Nop"] and not fl & F["EvalV"] for fl, *_ in prog)  # the read-after-write guard


@pytest.mark.parametrize("count", [1, 2, 5, 32])
@pytest.mark.parametrize("tip_children", [0, 1, 2])
def test_batched_root_evaluations_touch_nothing(count, tip_children):
    """the program rdk_root_loglikelihood_multi records (model_t::probe batches): `count`
    evaluations of ONE root operation, each with its own pair of P slots and its own result slot, no
    store at all -- every evaluation is the root term of its candidate and memory is left as it was,
    whether the root's children are inner CLVs, a tip and an inner CLV, or two tips"""
    rng = random.Random(100 + count)
    tips = 6
    ops, n_clv, n_sc = random_tree_ops(rng, tips, with_eval=False)
    inner = [o[0] for o in ops]
    if tip_children == 0:
        c1, c2 = inner[-1], inner[-2]
    elif tip_children == 1:
        c1, c2 = 2, inner[-1]
    else:
        c1, c2 = 0, 3
    sc_of = {o[0]: o[1] for o in ops}
    root, root_sc = n_clv, n_sc
    batch = [(root, root_sc, c1, c2, sc_of.get(c1, -1), sc_of.get(c2, -1), 100 + 2 * b, 101 + 2 * b, R_EVAL, b)
             for b in range(count)]
    ref, low = Mem(tips, n_clv + 1, n_sc + 1), Mem(tips, n_clv + 1, n_sc + 1)
    for mem in (ref, low):  # the traversal ran earlier: its CLVs are resident
        run_reference(mem, ops)
    before = (dict(low.clv), dict(low.sc))
    prog, _ = lower(tips, batch)
    want, got = run_reference(ref, batch), run_lowered(low, prog)
    assert got == want and len(got) == count
    assert (low.clv, low.sc) == before, "a batch of evaluations stores nothing"
    assert not any(fl & (F["Write"] | F["WriteS"]) for fl, *_ in prog)


# ---- subtree groups (rdk_partition_set_subtree_groups) ------------------------------------------------
def lower_grouped(tips, ops, n_groups, cap, discard=False):
    L = capi.load_engine()
    L.rdk_debug_lower_grouped.restype = C.c_int
    n = len(ops)
    arr = np.ascontiguousarray(np.array(ops, dtype=np.int64).astype(np.int32).reshape(n, 10))
    out = np.zeros((4 * n + 16, 10), dtype=np.int32)
    goff = np.zeros(n_groups + 1, dtype=np.uint32)
    ng, nt = C.c_uint(0), C.c_uint(0)
    used = L.rdk_debug_lower_grouped(
        C.c_uint(tips), C.c_uint(n), arr.ctypes.data_as(C.POINTER(C.c_int)), C.c_uint(n_groups), C.c_uint(cap),
        C.c_int(1 if discard else 0), out.ctypes.data_as(C.POINTER(C.c_int)), C.c_uint(out.shape[0]),
        goff.ctypes.data_as(C.POINTER(C.c_uint)), C.byref(ng), C.byref(nt))
    assert used >= 0
    if used == 0:
        return 0, [], [], []
    rows = [tuple(int(x) & 0xFFFFFFFF if j in (1, 3, 5) else int(x) for j, x in enumerate(row)) for row in out[:nt.value]]
    return used, rows[:ng.value], [int(x) for x in goff[:used + 1]], rows[ng.value:]


def post_order_ops(rng, tips, shape="random"):
    """post-order traversal of ONE rooted binary tree (what rooted_tree_t::generate_operations emits,
    reference src/tree.cpp:364-413), root operation evaluated"""
    class N:
        pass
    pool = []
    for t in range(tips):
        n = N()
        n.clv, n.sc, n.kids, n.pm = t, -1, None, t
        pool.append(n)
    clv, sc = tips, 0
    while len(pool) > 1:
        if shape == "caterpillar":
            i, j = len(pool) - 1, 0
        else:
            i, j = rng.sample(range(len(pool)), 2)
        a, b = pool[i], pool[j]
        for x in sorted((i, j), reverse=True):
            pool.pop(x)
        n = N()
        n.clv, n.sc, n.kids, n.pm = clv, sc, (a, b), clv
        clv, sc = clv + 1, sc + 1
        pool.append(n)
    ops = []

    def walk(n):
        if n.kids:
            walk(n.kids[0])
            walk(n.kids[1])
            ops.append((n.clv, n.sc, n.kids[0].clv, n.kids[1].clv, n.kids[0].sc, n.kids[1].sc,
                        n.kids[0].pm, n.kids[1].pm, R_WRITE, 0))
    import sys
    sys.setrecursionlimit(10000)
    walk(pool[0])
    ops[-1] = ops[-1][:8] + (R_WRITE | R_EVAL, 0)
    return ops, clv, sc


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("discard", [False, True])
def test_subtree_groups_compute_the_traversal(seed, discard):
    rng = random.Random(300 + seed)
    tips = rng.choice([6, 13, 40, 120, 500])
    ops, n_clv, n_sc = post_order_ops(rng, tips)
    n_groups = rng.choice([2, 3, 4, 8])
    cap = -(-len(ops) // (n_groups * rng.choice([1, 2])))
    used, groups, goff, join = lower_grouped(tips, ops, n_groups, cap, discard)
    ref = Mem(tips, n_clv, n_sc)
    want = run_reference(ref, ops)
    if used == 0:
        assert tips <= 13  # too small to have two subtrees under the cap
        return
    assert 2 <= used <= n_groups and goff[0] == 0 and goff[-1] == len(groups)
    assert all(goff[g] < goff[g + 1] for g in range(used))
    # groups are independent: run them in REVERSE order, each with unknown registers, then the join
    low = Mem(tips, n_clv, n_sc)
    got = {}
    for g in reversed(range(used)):
        got.update(run_lowered(low, groups[goff[g]:goff[g + 1]]))
    assert got == {}  # evaluations belong to the joining program
    got.update(run_lowered(low, join))
    assert got == want
    if not discard:
        assert ref.clv == low.clv and ref.sc == low.sc
    else:
        # a lazily materialised evaluation keeps the stores a group's own instructions or the joining
        # program read back; no group writes a buffer another group reads or writes
        written = [set(r[1] for r in groups[goff[g]:goff[g + 1]] if r[0] & F["Write"]) for g in range(used)]
        for g in range(used):
            for h in range(g + 1, used):
                assert not (written[g] & written[h])
    # no group reads what another group (or the join) writes
    produced_by = {}
    for g in range(used):
        for r in groups[goff[g]:goff[g + 1]]:
            if r[1] != NONE:
                produced_by[r[1]] = g
    for g in range(used):
        for (fl, parent, pscale, c1, s1, c2, pm1, pm2, slot, s2) in groups[goff[g]:goff[g + 1]]:
            if not (fl & (F["Tip1"] | F["Nop"])) and c1 in produced_by:
                assert produced_by[c1] == g
            if (fl & F["LoadV2"]) and c2 in produced_by:
                assert produced_by[c2] == g
    # balance: the longest group is not much longer than its share (random trees)
    longest = max(goff[g + 1] - goff[g] for g in range(used))
    assert longest + len(join) < len(ops) or tips <= 13


def test_subtree_groups_leave_other_programs_alone():
    rng = random.Random(9)
    # a caterpillar has no two disjoint subtrees worth dealing: everything joins
    ops, n_clv, n_sc = post_order_ops(rng, 64, shape="caterpillar")
    used, *_ = lower_grouped(64, ops, 4, len(ops) // 4)
    assert used == 0
    # a shared child (both parents read CLV 8) is not a forest
    tips = 8
    shared = [(8, 0, 0, 1, -1, -1, 0, 1, R_WRITE, 0), (9, 1, 8, 2, 0, -1, 2, 3, R_WRITE, 0),
              (10, 2, 8, 3, 0, -1, 4, 5, R_WRITE, 0), (11, 3, 9, 10, 1, 2, 6, 7, R_WRITE | R_EVAL, 0)]
    assert lower_grouped(tips, shared, 2, 2)[0] == 0
    # a buffer written twice / read before it is written keeps its order
    twice = [(8, 0, 0, 1, -1, -1, 0, 1, R_WRITE, 0), (9, 1, 2, 3, -1, -1, 2, 3, R_WRITE, 0),
             (8, 0, 4, 5, -1, -1, 4, 5, R_WRITE, 0), (10, 2, 8, 9, 0, 1, 6, 7, R_WRITE | R_EVAL, 0)]
    assert lower_grouped(tips, twice, 2, 2)[0] == 0
    war = [(9, 1, 8, 0, -1, -1, 0, 1, R_WRITE, 0), (8, 0, 1, 2, -1, -1, 2, 3, R_WRITE, 0),
           (10, 2, 8, 9, 0, 1, 6, 7, R_WRITE | R_EVAL, 0)]
    assert lower_grouped(tips, war, 2, 2)[0] == 0
    # the sequential random joins of the first test are forests too (not post-orders): same values
    for seed in range(6):
        r2 = random.Random(seed)
        t2 = r2.choice([17, 40, 90])
        ops2, c2, s2 = random_tree_ops(r2, t2)
        used, groups, goff, join = lower_grouped(t2, ops2, 3, -(-len(ops2) // 3))
        ref, low = Mem(t2, c2, s2), Mem(t2, c2, s2)
        want = run_reference(ref, ops2)
        if used == 0:
            continue
        got = {}
        for g in range(used):
            got.update(run_lowered(low, groups[goff[g]:goff[g + 1]]))
        got.update(run_lowered(low, join))
        assert got == want and ref.clv == low.clv and ref.sc == low.sc


def choose_groups(tips, ops, sites, cats, sm_count=148):
    L = capi.load_engine()
    L.rdk_debug_choose_subtree_groups.restype = C.c_int
    n = len(ops)
    arr = np.ascontiguousarray(np.array(ops, dtype=np.int64).astype(np.int32).reshape(n, 10))
    longest, n_join = C.c_uint(0), C.c_uint(0)
    g = L.rdk_debug_choose_subtree_groups(C.c_uint(tips), C.c_uint(n), arr.ctypes.data_as(C.POINTER(C.c_int)),
                                          C.c_uint(sites), C.c_uint(cats), C.c_int(sm_count), C.byref(longest),
                                          C.byref(n_join))
    return g, longest.value, n_join.value


def test_the_cost_model_groups_small_shards_only():
    """rdk_abi.cu choose_subtree_groups on a 148-SM device: the shard one GPU holds when
    cfg2 runs on 8 / 4 GPUs and the cfg3 shard are grouped, cfg2 on one or two GPUs and the cfg5 shard
    (throughput bound) stay one program in array order, and so does anything short"""
    rng = random.Random(1)
    ops500, *_ = post_order_ops(rng, 500)
    g, longest, n_join = choose_groups(500, ops500, 12544, 4)
    assert g == 4 and longest <= 140 and n_join <= 24 and longest * g + n_join >= 499
    g, longest, n_join = choose_groups(500, ops500, 25088, 4)
    assert g in (2, 3) and longest <= 280
    assert choose_groups(500, ops500, 50176, 4)[0] == 0
    assert choose_groups(500, ops500, 100000, 4)[0] == 0
    ops2000, *_ = post_order_ops(rng, 2000)
    g, longest, n_join = choose_groups(2000, ops2000, 62464, 4)
    assert g >= 4 and longest <= 2 * 1999 // g
    ops10k, *_ = post_order_ops(rng, 10000)
    assert choose_groups(10000, ops10k, 125184, 4)[0] == 0
    # too short to be worth a second launch; a caterpillar has nothing to deal
    ops30, *_ = post_order_ops(rng, 30)
    assert choose_groups(30, ops30, 1000, 4)[0] == 0
    cat, *_ = post_order_ops(rng, 400, shape="caterpillar")
    assert choose_groups(400, cat, 12544, 4)[0] == 0
    # any category count: the device works on the next power of two
    assert choose_groups(500, ops500, 12544, 3)[0] == 4
