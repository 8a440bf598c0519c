"""CPU: the compact layout the vote kernel for small trees walks (csrc/abi.cu build_compact, through the host-only
hook rfmc_debug_compact_layout) -- every row ends on the same payload as the walk over the 16-byte nodes, in
exactly ``depth[tree]`` branch-free steps; stages respect their record and tree limits."""
import ctypes as C

import numpy as np
import pytest

from random_forest_mc_b200 import engine
from random_forest_mc_b200.flat import forest_tables_from_dicts
from tests.helpers import golden_names, load_golden

LEAF, CAT = 0x80000000, 0x40000000


def compact_layout(tables):
    lib = engine.load_library()
    fn = lib.rfmc_debug_compact_layout
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 2 + [C.c_int32] + [C.c_void_p] * 9
    nodes = np.ascontiguousarray(tables.pnodes)
    roots = np.ascontiguousarray(tables.roots, dtype=np.int64)
    n_rec, n_st, has4 = C.c_int64(0), C.c_int32(0), C.c_int32(0)
    T = len(roots)
    if fn(nodes.ctypes.data, roots.ctypes.data, T, C.addressof(n_rec), C.addressof(n_st), None, None, None, None, None, None, C.addressof(has4)):
        return None
    rec = np.zeros((n_rec.value, 2), dtype=np.uint32)
    rec4 = np.zeros(n_rec.value, dtype=np.uint32)
    pay = np.zeros(n_rec.value, dtype=np.uint32)
    tree_at = np.zeros(T + 1, dtype=np.uint32)
    depth = np.zeros(T, dtype=np.uint8)
    stage = np.zeros(n_st.value + 1, dtype=np.uint32)
    assert fn(nodes.ctypes.data, roots.ctypes.data, T, None, None, rec.ctypes.data, pay.ctypes.data, tree_at.ctypes.data,
              depth.ctypes.data, stage.ctypes.data, rec4.ctypes.data, None) == 0
    LAST_REC4[0] = rec4 if has4.value else None
    return rec, pay, tree_at, depth, stage


LAST_REC4 = [None]  # the 4-byte records of the last layout asked for (None: the forest is not eligible for them)


def walk_rec4(rec4, layout, t, row):
    """The kernel's walk over 4-byte records {feat | delta << 8 | numeric << 15 | lo << 16}: next = own index + delta."""
    _, pay, tree_at, depth, _ = layout
    base, idx = int(tree_at[t]), 0
    for _ in range(int(depth[t])):
        r = int(rec4[base + idx])
        d = (int(row[r & 0xFF]) - (r >> 16)) & 0xFFFFFFFF
        idx = idx + ((r >> 8) & 0x7F) + (1 if d > ((r & 0x8000) << 1) else 0)
    assert int(rec4[base + idx]) == 0x8000, "the walk must end on a leaf record"
    return int(pay[base + idx])


def walk_pnodes(tables, t, row):
    ff, thr, child = tables.pnodes["feat_flags"], tables.pnodes["thr"], tables.pnodes["child"]
    i = int(tables.roots[t])
    while not (int(ff[i]) & LEAF):
        c = int(row[int(ff[i]) & 0xFFFF])
        go = (c == int(thr[i])) if (int(ff[i]) & CAT) else (c >= int(thr[i]))
        i = int(child[i]) + (0 if go else 1)
    return int(thr[i])


def walk_compact(layout, t, row):
    rec, pay, tree_at, depth, _ = layout
    base, idx = int(tree_at[t]), 0
    for _ in range(int(depth[t])):
        x, y = int(rec[base + idx, 0]), int(rec[base + idx, 1])
        d = (int(row[x & 0xFFFF]) - (x >> 16)) & 0xFFFFFFFF  # the kernel's unsigned 32-bit subtraction
        idx = (y >> 16) + (0 if d <= (y & 0xFFFF) else 1)
    x, y = int(rec[base + idx, 0]), int(rec[base + idx, 1])
    assert x == 0 and (y & 0xFFFF) == 0xFFFF and (y >> 16) == idx, "the walk must end on a self-looping leaf record"
    return int(pay[base + idx])


@pytest.mark.parametrize("name", golden_names())
def test_compact_walk_equals_node_walk(name):
    g = load_golden(name)
    tables = forest_tables_from_dicts(g["forest"], [float(s) for s in g["survived_scores"]], g["class_vals"], g["feature_cols"])
    layout = compact_layout(tables)
    assert layout is not None, "golden forests are small: they must be eligible"
    rec, pay, tree_at, depth, stage = layout
    T = tables.n_trees
    assert tree_at[0] == 0 and tree_at[-1] == len(pay) == len(tables.pnodes)
    assert stage[0] == 0 and stage[-1] == T and np.all(np.diff(stage.astype(np.int64)) > 0)
    for s in range(len(stage) - 1):
        assert int(tree_at[stage[s + 1]]) - int(tree_at[stage[s]]) <= 1024 and int(stage[s + 1]) - int(stage[s]) <= 64
    rng = np.random.default_rng(7)
    hi = [1 + (len(t) if t is not None else 0) for t in tables.thresholds]
    for j, cats in enumerate(tables.categories):
        if cats:
            hi[j] = 1 + len(cats)
    rec4 = LAST_REC4[0]  # None for a forest with a level wider than 127 nodes (mixed_big)
    assert rec4 is not None or name not in ("titanic_c1", "mixed_default", "ties_default")
    for _ in range(300):
        row = np.array([rng.integers(0, h + 1) for h in hi], dtype=np.int64)  # code 0 = missing / unseen
        for t in range(T):
            want = walk_pnodes(tables, t, row)
            assert walk_compact(layout, t, row) == want
            if rec4 is not None:
                assert walk_rec4(rec4, layout, t, row) == want


def test_big_trees_are_not_eligible_and_many_small_trees_split_into_stages():
    # a chain of 600 splits = 1,201 nodes > one stage
    from random_forest_mc_b200.flat import PNODE_DTYPE

    n = 1201
    nodes = np.zeros(n, dtype=PNODE_DTYPE)  # split k: children at (nxt, nxt + 1), the '>=' child is the next split
    at = 0
    nxt = 1
    for k in range(600):
        nodes[at]["feat_flags"], nodes[at]["thr"], nodes[at]["child"] = 0, 1, nxt
        nodes[nxt + 1]["feat_flags"], nodes[nxt + 1]["thr"] = LEAF, 0  # '<' child: a leaf
        at = nxt  # '>=' child: the next split
        nxt += 2
    nodes[at]["feat_flags"], nodes[at]["thr"] = LEAF, 0

    class T_:  # minimal stand-in for ForestTables
        pnodes, roots = nodes, np.array([0], dtype=np.int64)

    assert compact_layout(T_) is None
    # 200 one-split trees: 600 records, more than 64 trees per stage -> at least four stages
    small = np.zeros(600, dtype=PNODE_DTYPE)
    for t in range(200):
        small[3 * t]["feat_flags"], small[3 * t]["thr"], small[3 * t]["child"] = 0, 1, 3 * t + 1
        small[3 * t + 1]["feat_flags"], small[3 * t + 1]["thr"] = LEAF, 2 * t
        small[3 * t + 2]["feat_flags"], small[3 * t + 2]["thr"] = LEAF, 2 * t + 1

    class S_:
        pnodes, roots = small, np.arange(0, 600, 3, dtype=np.int64)

    rec, pay, tree_at, depth, stage = compact_layout(S_)
    assert len(stage) - 1 >= 4 and np.all(depth == 1)
    for t in range(200):
        assert walk_compact((rec, pay, tree_at, depth, stage), t, np.array([1])) == 2 * t
        assert walk_compact((rec, pay, tree_at, depth, stage), t, np.array([0])) == 2 * t + 1


def test_wide_forests_keep_eight_byte_records():
    """A feature id of 256 or more, or a '>=' child more than 127 records behind its parent, rules out 4-byte records."""
    from random_forest_mc_b200.flat import PNODE_DTYPE

    one = np.zeros(3, dtype=PNODE_DTYPE)
    one[0]["feat_flags"], one[0]["thr"], one[0]["child"] = 300, 1, 1
    one[1]["feat_flags"], one[2]["feat_flags"] = LEAF, LEAF

    class W_:
        pnodes, roots = one, np.array([0], dtype=np.int64)

    assert compact_layout(W_) is not None and LAST_REC4[0] is None
    # a complete tree of depth 8: level 7 has 128 nodes, so the children of its first node are 128+ records away
    n = 2 ** 9 - 1
    full = np.zeros(n, dtype=PNODE_DTYPE)
    for i in range(n):
        if 2 * i + 2 < n:
            full[i]["feat_flags"], full[i]["thr"], full[i]["child"] = 0, 1, 2 * i + 1
        else:
            full[i]["feat_flags"], full[i]["thr"] = LEAF, i

    class F_:
        pnodes, roots = full, np.array([0], dtype=np.int64)

    layout = compact_layout(F_)
    assert layout is not None and LAST_REC4[0] is None
    for code in (0, 1):
        assert walk_compact(layout, 0, np.array([code])) == walk_pnodes(F_, 0, np.array([code]))
