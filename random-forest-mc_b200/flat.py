"""SoA node tables <-> the reference's nested tree dicts.

The engine's native tree is a flat node table (what the CUDA grower writes and what the traversal
kernel reads).  The reference's native tree is a nested dict (tree.py:44-99, produced at
model.py:282-291 and :226-229).  This module converts between the two losslessly, including dict
key order and Python/numpy value types:

* split node  ``{feat: {"split": {"feat_type", "split_val", ">=", "<"}}}``
  - >2-row numeric node: ``split_val`` is a Python float (``float(np.quantile(..))``, model.py:237)
  - 2-row node: ``split_val`` is the native element of the column array (model.py:262)
  - categorical node: the native element of ``np.unique``'s output (model.py:250)
* leaf ``{"leaf": {label: count/total ...present labels, lexicographic...}, "depth": "k#"}``
"""

from __future__ import annotations

import gc

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .encode import KIND_CATEGORICAL, EncodedDataset

FLAG_CATEGORICAL = 1
FLAG_TWO_ROW = 2

# Mirrors ``struct rfmc_node`` in include/rfmc_b200.h (24 bytes).
NODE_DTYPE = np.dtype(
    [("feat", "<i4"), ("thr", "<u4"), ("child", "<i4"), ("flags", "<u4"), ("sval", "<f8")], align=True
)
assert NODE_DTYPE.itemsize == 24


@dataclass
class TreeSoA:
    """One grown tree: ``nodes`` (NODE_DTYPE, BFS order, root = 0, '<' child = '>=' child + 1) and
    ``counts[node][class]`` (sorted-label order; meaningful for leaves)."""

    nodes: np.ndarray
    counts: np.ndarray

    @property
    def n_nodes(self) -> int:
        return len(self.nodes)


def soa_to_dict(tree: TreeSoA, enc: EncodedDataset, type_of_cols: Optional[dict] = None) -> dict:
    """Materialise the nested dict the reference would have built for this tree.

    Nodes are in BFS order (a child's index is larger than its parent's), so the dicts are built bottom-up in
    one reverse pass over plain Python lists; everything that can be computed for all nodes at once (depths,
    leaf probabilities, split values) is, because a forest grown to purity has millions of nodes and
    ``model2dict()`` / ``tree.data`` pay this per node."""
    nodes, counts = tree.nodes, tree.counts
    n = len(nodes)
    feat_a = nodes["feat"]
    feat, thr, child, flags, sval = (nodes[k].tolist() for k in ("feat", "thr", "child", "flags", "sval"))
    labels = enc.class_sorted
    cols = enc.columns
    names = [c.name for c in cols]
    if type_of_cols is not None:
        kinds = [type_of_cols[c.name] for c in cols]
    else:
        kinds = ["categorical" if c.kind == KIND_CATEGORICAL else "numeric" for c in cols]
    uniq = {}  # per used column: its value table as a list of numpy scalars (what uniq[k] returns)

    # depth of every node (root = 1): parents come first in BFS order
    depth = [1] * n
    for i in range(n):
        if feat[i] >= 0:
            c = child[i]
            depth[c] = depth[c + 1] = depth[i] + 1
    dstr = {}

    # leaf payloads {label: count / total} in sorted-label order, np.float64 values as genLeaf's (model.py:222-229)
    leaf_idx = np.flatnonzero(feat_a < 0)
    cnt = np.asarray(counts)[leaf_idx]
    tot = cnt.sum(axis=1)
    rr, cc = np.nonzero(cnt)  # row-major: per leaf, classes ascending
    vals = list(cnt[rr, cc].astype(np.int64) / tot[rr]) if len(rr) else []
    lab_of = [labels[c] for c in cc.tolist()]
    first = np.searchsorted(rr, np.arange(len(leaf_idx) + 1)).tolist()
    leaf_pos = dict(zip(leaf_idx.tolist(), range(len(leaf_idx))))

    objs = [None] * n
    special = FLAG_CATEGORICAL | FLAG_TWO_ROW
    # tens of thousands of small dicts and no reference cycle among them: the cyclic collector would only
    # re-traverse the growing forest again and again
    gc_was_on = gc.isenabled()
    gc.disable()
    try:
        _build_dicts(n, feat, thr, child, flags, sval, depth, dstr, leaf_pos, first, lab_of, vals, uniq, cols, names, kinds, objs, special)
    finally:
        if gc_was_on:
            gc.enable()
    return objs[0]


def _build_dicts(n, feat, thr, child, flags, sval, depth, dstr, leaf_pos, first, lab_of, vals, uniq, cols, names, kinds, objs, special):
    for i in range(n - 1, -1, -1):
        f = feat[i]
        if f < 0:
            k = leaf_pos[i]
            a, b = first[k], first[k + 1]
            leaf = {lab_of[a]: vals[a]} if b - a == 1 else dict(zip(lab_of[a:b], vals[a:b]))
            d = depth[i]
            ds = dstr.get(d)
            if ds is None:
                ds = dstr[d] = f"{d}#"
            objs[i] = {"leaf": leaf, "depth": ds}
            continue
        if flags[i] & special:
            u = uniq.get(f)
            if u is None:
                u = uniq[f] = list(cols[f].uniq)
            value = u[thr[i] - 1]
        else:
            value = sval[i]
        c = child[i]
        objs[i] = {names[f]: {"split": {"feat_type": kinds[f], "split_val": value, ">=": objs[c], "<": objs[c + 1]}}}


def used_feature_ids(tree: TreeSoA) -> List[int]:
    f = tree.nodes["feat"]
    return sorted(set(int(x) for x in np.unique(f[f >= 0])))


_RE_QUOTED_KEY = __import__("re").compile(r"[\w\s]+")


def reference_used_features(feature_cols, split_names, leaf_labels, has_split: bool) -> list:
    """What the reference's ``tree2feats`` (model.py:413-417) reports for a tree: it greps ``str(tree)`` for
    quoted keys (``'[\\w\\s]+':``) and intersects them with ``"'{f}':"`` of every feature column.  So a feature counts
    as used when its name -- formatted into a string -- equals ANY string key of the nested dict made of word
    characters and blanks: a feature it really splits on, but also the structural keys ('leaf', 'depth' in every tree;
    'split', 'feat_type', 'split_val' when there is a split) and the class labels of its leaves; a feature whose name is
    not a string, or holds other characters, is never found.  Returned in ``feature_cols`` order (the reference's order
    is that of a hash set)."""
    keys = {"leaf", "depth"}
    if has_split:
        keys.update(("split", "feat_type", "split_val"))
    keys.update(k for k in split_names if isinstance(k, str))
    keys.update(k for k in leaf_labels if isinstance(k, str))
    keys = {k for k in keys if _RE_QUOTED_KEY.fullmatch(k)}
    return [f"{f}" for f in feature_cols if f"{f}" in keys]


def reference_used_features_of_dict(feature_cols, tree: dict) -> list:
    split_names, labels, todo = set(), set(), [tree]
    while todo:
        node = todo.pop()
        name = next(iter(node))
        if name == "leaf":
            labels.update(node["leaf"].keys())
        else:
            split_names.add(name)
            todo.append(node[name]["split"][">="])
            todo.append(node[name]["split"]["<"])
    return reference_used_features(feature_cols, split_names, labels, bool(split_names))


def reference_used_features_of_soa(feature_cols, tree: TreeSoA, class_sorted) -> list:
    feat = tree.nodes["feat"]
    leaf = feat < 0
    present = np.flatnonzero(tree.counts[leaf].sum(axis=0) > 0)
    ids = used_feature_ids(tree)
    return reference_used_features(feature_cols, [feature_cols[j] for j in ids], [class_sorted[k] for k in present], len(ids) > 0)


# --------------------------------------------------------------------------------------------
# Forest flattening for the voting kernel
# --------------------------------------------------------------------------------------------
PNODE_DTYPE = np.dtype([("feat_flags", "<u4"), ("thr", "<u4"), ("child", "<u4"), ("reserved", "<u4")])
PNODE_LEAF = 0x80000000
PNODE_CATEGORICAL = 0x40000000


@dataclass
class ForestTables:
    """A forest flattened for rfmc_forest_create, plus the per-feature code books that map raw
    frame values into the code space its thresholds live in.

    Code space (last paragraph): for a numeric feature the sorted distinct split
    values t_0 < t_1 < ... of the whole forest; ``code(x) = #{k : t_k <= x}`` so that
    ``x >= t_k  <=>  code(x) >= k + 1``; NaN -> 0 (fails every test -> '<', tree.py:177-181).
    For a categorical feature every distinct split value gets an id >= 1; anything else -> 0."""

    feature_cols: List[str]
    class_vals: List
    pnodes: np.ndarray
    roots: np.ndarray  # int64 [n_trees]
    vote: np.ndarray  # uint8 [n_payloads] class_vals index
    vote_absent: np.ndarray
    probs: np.ndarray  # float64 [n_payloads, C]
    probs_absent: np.ndarray
    scores: np.ndarray  # float64 [n_trees]
    score_fsum: float
    thresholds: List[Optional[np.ndarray]]  # per feature (numeric)
    categories: List[Optional[dict]]  # per feature (categorical): value -> id
    width: int = 1
    stride: int = 0
    payload_leaf: Optional[list] = None  # dict path only: the leaf dict behind every payload

    @property
    def n_classes(self) -> int:
        return len(self.class_vals)

    @property
    def n_features(self) -> int:
        return len(self.feature_cols)

    @property
    def n_trees(self) -> int:
        return len(self.roots)


def leaf_payload(leaf: dict, class_vals, cidx):
    """(vote, vote_absent, probs, probs_absent) of one leaf dict.
    tree.py:185-212 for the absent variant; forest.py:200-202 for the votes."""
    from math import fsum

    C = len(class_vals)
    probs = [0.0] * C
    best, bestp = -1, None
    for lab, p in leaf.items():  # key order of the leaf dict decides ties
        k = cidx[lab]
        probs[k] = float(p)
        if bestp is None or p > bestp:
            best, bestp = k, p
    out = [0 + probs[k] for k in range(C)]
    out = [v / 1 for v in out]
    total = fsum(out)
    pa = [v / total for v in out]
    besta, bestpa = 0, pa[0]
    for k in range(1, C):
        if pa[k] > bestpa:
            besta, bestpa = k, pa[k]
    return best, besta, tuple(probs), tuple(pa)


def weak_float32(v) -> float:
    """A Python-number split value as a float32 cell sees it: numpy compares a np.float32 with a Python
    scalar in float32 (NEP 50), i.e. against the split value rounded to float32."""
    with np.errstate(over="ignore"):
        return float(np.float32(v))


def forest_tables_from_dicts(trees, scores, class_vals, feature_cols, f32_feats=frozenset()) -> ForestTables:
    """Flatten nested tree dicts (fitted here, merged, or loaded through dict2model).  ``f32_feats``:
    features whose cells the frame hands over as np.float32 scalars (see forest._float32_features)."""
    from math import fsum

    def eff(v, j):
        return weak_float32(v) if (j in f32_feats and type(v) in (float, int)) else float(v)

    from .encode import odd_word_stride

    fidx = {name: j for j, name in enumerate(feature_cols)}
    cidx = {c: k for k, c in enumerate(class_vals)}
    nfeat = len(feature_cols)
    if nfeat > 0xFFFF:
        raise NotImplementedError("more than 65535 feature columns")
    num_vals = [set() for _ in range(nfeat)]
    cat_vals = [dict() for _ in range(nfeat)]
    # pass 1: code books
    for tree in trees:
        todo = [tree]
        while todo:
            node = todo.pop()
            name = next(iter(node))
            if name == "leaf":
                continue
            s = node[name]["split"]
            j = fidx[name]
            if s["feat_type"] != "numeric":
                if s["split_val"] not in cat_vals[j]:
                    cat_vals[j][s["split_val"]] = len(cat_vals[j]) + 1
            else:
                num_vals[j].add(eff(s["split_val"], j))
            todo.append(s[">="])
            todo.append(s["<"])
    thresholds, categories = [], []
    top = 1
    for j in range(nfeat):
        if num_vals[j] and cat_vals[j]:
            raise NotImplementedError(f"feature {feature_cols[j]!r} is split both as numeric and as categorical")
        t = np.array(sorted(v for v in num_vals[j] if v == v), dtype=np.float64) if num_vals[j] else None
        thresholds.append(t)
        categories.append(cat_vals[j] if cat_vals[j] else None)
        top = max(top, (len(t) if t is not None else 0), len(cat_vals[j]))
    width = 1 if top <= 0xFF else (2 if top <= 0xFFFF else 4)
    # pass 2: nodes (children adjacent: '<' = '>=' + 1) and de-duplicated leaf payloads
    feat_flags, thr, child, roots = [], [], [], []
    payload_ids = {}
    votes, votes_a, probs, probs_a, leaf_dicts = [], [], [], [], []
    for tree in trees:
        root = len(feat_flags)
        roots.append(root)
        feat_flags.append(0), thr.append(0), child.append(0)
        todo = [(tree, root)]
        while todo:
            node, at = todo.pop()
            name = next(iter(node))
            if name == "leaf":
                key = tuple(node["leaf"].items())
                pid = payload_ids.get(key)
                if pid is None:
                    v, va, p, pa = leaf_payload(node["leaf"], class_vals, cidx)
                    pid = payload_ids[key] = len(votes)
                    votes.append(v), votes_a.append(va), probs.append(p), probs_a.append(pa)
                    leaf_dicts.append(node["leaf"])
                feat_flags[at] = PNODE_LEAF
                thr[at] = pid
                continue
            s = node[name]["split"]
            j = fidx[name]
            if s["feat_type"] != "numeric":
                feat_flags[at] = j | PNODE_CATEGORICAL
                thr[at] = cat_vals[j][s["split_val"]]
            else:
                feat_flags[at] = j
                sv = eff(s["split_val"], j)
                # a NaN split value matches nothing: an unreachable code sends every row to '<'
                thr[at] = int(np.searchsorted(thresholds[j], sv, "left")) + 1 if sv == sv else 0xFFFFFFFF
            c = len(feat_flags)
            child[at] = c
            feat_flags.extend((0, 0)), thr.extend((0, 0)), child.extend((0, 0))
            todo.append((s["<"], c + 1))
            todo.append((s[">="], c))
    pn = np.zeros(len(feat_flags), dtype=PNODE_DTYPE)
    pn["feat_flags"] = np.array(feat_flags, dtype=np.uint32)
    pn["thr"] = np.array(thr, dtype=np.uint32)
    pn["child"] = np.array(child, dtype=np.uint32)
    ncls = len(class_vals)
    return ForestTables(
        feature_cols=list(feature_cols),
        class_vals=list(class_vals),
        pnodes=pn,
        roots=np.array(roots, dtype=np.int64),
        vote=np.array(votes, dtype=np.uint8),
        vote_absent=np.array(votes_a, dtype=np.uint8),
        probs=np.array(probs, dtype=np.float64).reshape(len(votes), ncls),
        probs_absent=np.array(probs_a, dtype=np.float64).reshape(len(votes), ncls),
        scores=np.array([float(s) for s in scores], dtype=np.float64),
        score_fsum=fsum(float(s) for s in scores),
        thresholds=thresholds,
        categories=categories,
        width=width,
        stride=odd_word_stride(nfeat, width),
        payload_leaf=leaf_dicts,
    )


def encode_frame(tables: ForestTables, frame):
    """Raw predict frame -> (codes [n, stride], absent [n_features] uint8).

    Mirrors what ``ds.iterrows()`` + ``row[node]`` hand to the comparisons of tree.py:176-180:
    a column missing from the frame is *absent* (tree.py:166-175); a NaN / None / unseen cell is
    code 0.  Numeric values are compared as float64 (exact for float32 and |int| < 2**53)."""
    import pandas as pd

    n = len(frame)
    dt = {1: np.uint8, 2: np.uint16, 4: np.uint32}[tables.width]
    codes = np.zeros((n, tables.stride), dtype=dt)
    absent = np.zeros(tables.n_features, dtype=np.uint8)
    have = set(frame.columns)
    for j, name in enumerate(tables.feature_cols):
        t, cats = tables.thresholds[j], tables.categories[j]
        if t is None and cats is None:
            continue  # never tested by any tree
        if name not in have:
            absent[j] = 1
            continue
        col = frame[name]
        if t is not None:
            x = pd.to_numeric(col, errors="coerce").to_numpy(dtype=np.float64, na_value=np.nan)
            code = np.searchsorted(t, x, "right")
            code[np.isnan(x)] = 0
            codes[:, j] = code
        else:
            ids, uniques = col.factorize(use_na_sentinel=True)
            lut = np.zeros(len(uniques) + 1, dtype=np.int64)  # last slot: NaN sentinel (-1) -> 0
            for k, u in enumerate(uniques):
                try:
                    lut[k] = cats.get(u, 0)
                except TypeError:
                    lut[k] = 0
            codes[:, j] = lut[ids]
    return codes, absent


def forest_tables_from_soa(trees, scores, enc: EncodedDataset, class_vals, feature_cols, f32_feats=frozenset()) -> ForestTables:
    """Same tables as ``forest_tables_from_dicts`` but straight from the grower's SoA output,
    vectorised over all nodes of all trees (no nested dict is ever materialised)."""
    from math import fsum

    from .encode import odd_word_stride

    nfeat = len(feature_cols)
    sizes = np.array([t.n_nodes for t in trees], dtype=np.int64)
    roots = np.zeros(len(trees), dtype=np.int64)
    roots[1:] = np.cumsum(sizes)[:-1]
    nodes = np.concatenate([t.nodes for t in trees])
    counts = np.concatenate([t.counts for t in trees])
    feat = nodes["feat"].astype(np.int64)
    flags = nodes["flags"]
    thr_ds = nodes["thr"].astype(np.int64)
    is_leaf = feat < 0
    is_cat = (flags & FLAG_CATEGORICAL) != 0
    two_row = (flags & FLAG_TWO_ROW) != 0
    base = np.repeat(roots, sizes)

    ff = np.zeros(len(nodes), dtype=np.uint32)
    thr = np.zeros(len(nodes), dtype=np.uint32)
    child = np.where(is_leaf, 0, nodes["child"].astype(np.int64) + base).astype(np.uint32)
    thresholds, categories = [None] * nfeat, [None] * nfeat
    top = 1
    for j in np.unique(feat[~is_leaf]):
        j = int(j)
        col = enc.columns[j]
        m = (feat == j) & ~is_leaf
        if col.kind == KIND_CATEGORICAL:
            present = np.unique(thr_ds[m])
            thr[m] = (np.searchsorted(present, thr_ds[m]) + 1).astype(np.uint32)
            ff[m] = j | PNODE_CATEGORICAL
            categories[j] = {col.uniq[int(c) - 1]: k + 1 for k, c in enumerate(present)}
            top = max(top, len(present))
        else:
            sval = nodes["sval"][m]
            if j in f32_feats:  # Python-float split values (model.py:237) as a np.float32 cell sees them
                with np.errstate(over="ignore"):
                    sval = sval.astype(np.float32).astype(np.float64)
            vals = np.where(two_row[m], col.table[np.clip(thr_ds[m] - 1, 0, len(col.table) - 1)], sval)
            ok = vals == vals
            t = np.unique(vals[ok])
            code = np.full(len(vals), 0xFFFFFFFF, dtype=np.int64)
            code[ok] = np.searchsorted(t, vals[ok], "left") + 1
            thr[m] = code.astype(np.uint32)
            ff[m] = j
            thresholds[j] = t
            top = max(top, len(t))
    # leaf payloads, de-duplicated on the class-count vector
    leaf_rows = counts[is_leaf].astype(np.int64)
    uniq_rows, inverse = np.unique(leaf_rows, axis=0, return_inverse=True)
    inverse = np.asarray(inverse).reshape(-1)
    ff[is_leaf] = PNODE_LEAF
    thr[is_leaf] = inverse.astype(np.uint32)
    totals = uniq_rows.sum(axis=1)
    p_sorted = uniq_rows / totals[:, None]  # int64 / int64 -> correctly rounded, == count/total (model.py:227)
    s2c = np.asarray(enc.sorted_to_classvals, dtype=np.int64)
    ncls = len(class_vals)
    probs = np.zeros((len(uniq_rows), ncls), dtype=np.float64)
    probs[:, s2c] = p_sorted
    vote = s2c[np.argmax(uniq_rows, axis=1)].astype(np.uint8)  # first max in sorted-label order
    probs_absent = np.empty_like(probs)
    for r in range(len(probs)):
        row = [0 + float(v) for v in probs[r]]
        total = fsum(row)
        probs_absent[r] = [v / 1 / total for v in row]
    vote_absent = np.argmax(probs_absent, axis=1).astype(np.uint8)
    width = 1 if top <= 0xFF else (2 if top <= 0xFFFF else 4)
    pn = np.zeros(len(nodes), dtype=PNODE_DTYPE)
    pn["feat_flags"], pn["thr"], pn["child"] = ff, thr, child
    return ForestTables(
        feature_cols=list(feature_cols),
        class_vals=list(class_vals),
        pnodes=pn,
        roots=roots,
        vote=vote,
        vote_absent=vote_absent,
        probs=probs,
        probs_absent=probs_absent,
        scores=np.array([float(s) for s in scores], dtype=np.float64),
        score_fsum=fsum(float(s) for s in scores),
        thresholds=thresholds,
        categories=categories,
        width=width,
        stride=odd_word_stride(nfeat, width),
    )
