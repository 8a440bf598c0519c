"""CPU ORACLE for the RandomForestMC hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy / Python, what ysraell/random-forest-mc v1.4.0 computes on
the fit -> score -> keep/discard -> vote path.  It exists so that the CUDA engine can be checked
bit-for-bit against an independent CPU statement of the same algorithm.

Who may import this: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under ``random-forest-mc_b200/`` imports it and
the product has no CPU fallback.

Parity pin: the oracle is pinned against the *real* reference (imported from /root/reference in the
build container) by ``tests/golden/make_golden.py`` -> ``tests/golden/*.pkl.gz`` and checked in
``tests/test_oracle_golden.py`` (tree dicts incl. value types and key order, kept/discarded sets,
scores, labels and probabilities in all four voting modes, NaN cells, absent columns); in the build
container ``tests/test_oracle_live_reference.py`` also runs the real reference live on 30 randomized
cases and requires equality, RNG end states included.

Every function cites the reference lines it follows (paths relative to the reference repo root).
It deliberately issues the same numpy calls per node as the reference (np.quantile, np.unique,
np.argsort, np.random.choice ...), so that timing it is a fair stand-in for timing the reference
itself (the Python reference cannot travel to the GPU box).
"""

from __future__ import annotations

import random as _pyrandom
from collections import defaultdict
from dataclasses import dataclass, field
from math import fsum
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd

__all__ = [
    "OracleDataset",
    "ingest",
    "draw_rows",
    "draw_features",
    "split_node",
    "make_leaf",
    "grow",
    "walk",
    "tree_output",
    "top_class",
    "score_candidate",
    "select_tree",
    "fit",
    "forest_row",
    "forest_labels",
    "forest_probs",
]


# --------------------------------------------------------------------------------------------
# Dataset ingest                                           src/random_forest_mc/model.py:162-195
# --------------------------------------------------------------------------------------------
@dataclass
class OracleDataset:
    target_col: str
    feature_cols: List[str]
    numeric_cols: List[str]
    type_of_cols: Dict[str, str]
    class_vals: List[str]
    columns: Dict[str, np.ndarray]
    n_samples: int
    rows_per_class: int  # the reference's ``_N`` after the smallest-class cap
    batch_train_pclass: int
    batch_val_pclass: int
    min_feature: int
    max_feature: int
    temporal_features: bool = False
    max_depth: int = 1000
    min_samples_split: int = 1
    class_rows: Dict[str, np.ndarray] = field(default_factory=dict)


def ingest(
    frame: pd.DataFrame,
    target_col: str = "target",
    batch_train_pclass: int = 10,
    batch_val_pclass: int = 10,
    min_feature: Optional[int] = None,
    max_feature: Optional[int] = None,
    temporal_features: bool = False,
    max_depth: Optional[int] = None,
    min_samples_split: int = 1,
) -> OracleDataset:
    """model.py:162-195.  Mutates ``frame`` like the reference does (target -> str, :163)."""
    frame[target_col] = frame[target_col].astype(str)
    feats = [c for c in frame.columns if c != target_col]
    numeric = frame.select_dtypes([np.number]).columns.to_list()  # :165
    categorical = list(set(feats) - set(numeric))  # :166
    kinds = {c: "numeric" for c in numeric}
    kinds.update({c: "categorical" for c in categorical})
    lo = 2 if min_feature is None else min_feature  # :170-171
    hi = len(feats) if max_feature is None else max_feature  # :173-174
    clean = frame.dropna()  # :181
    classes = clean[target_col].unique().tolist()  # :184 first-appearance order
    temporal = bool(temporal_features)
    if temporal and not all(x.split("_")[-1].isnumeric() for x in feats):  # :186-188, forest.py:148
        temporal = False
    per_class = min(batch_train_pclass + batch_val_pclass, clean[target_col].value_counts().min())  # :190
    cols = {c: clean[c].to_numpy() for c in clean.columns}  # :194
    import sys

    return OracleDataset(
        target_col=target_col,
        feature_cols=feats,
        numeric_cols=numeric,
        type_of_cols=kinds,
        class_vals=classes,
        columns=cols,
        n_samples=len(clean),
        rows_per_class=int(per_class),
        batch_train_pclass=batch_train_pclass,
        batch_val_pclass=batch_val_pclass,
        min_feature=lo,
        max_feature=hi,
        temporal_features=temporal,
        max_depth=sys.getrecursionlimit() if max_depth is None else int(max_depth),  # :156
        min_samples_split=int(min_samples_split),
    )


# --------------------------------------------------------------------------------------------
# Host RNG draws                                          src/random_forest_mc/model.py:198-219
# --------------------------------------------------------------------------------------------
def draw_rows(ds: OracleDataset) -> Tuple[np.ndarray, np.ndarray]:
    """model.py:198-210: class-balanced draw without replacement from numpy's GLOBAL legacy RNG.

    The object-array compare ``target == val`` (:205) consumes no randomness, so caching the
    per-class row lists would not change the stream; the oracle recomputes them like the reference
    so its timing stays representative."""
    train: List[Any] = []
    val: List[Any] = []
    y = ds.columns[ds.target_col]
    everything = np.arange(ds.n_samples)
    for label in ds.class_vals:
        pool = everything[y == label]
        picked = np.random.choice(pool, ds.rows_per_class, replace=False)  # :206
        train.extend(picked[: ds.batch_train_pclass])
        val.extend(picked[ds.batch_train_pclass :])
    return np.array(train), np.array(val)


def draw_features(ds: OracleDataset) -> List[str]:
    """model.py:213-219: count from numpy's RNG, subset+order from CPython's ``random``."""
    pool = ds.feature_cols[:]
    k = np.random.randint(ds.min_feature, ds.max_feature + 1)
    chosen = _pyrandom.sample(pool, min(len(pool), k))
    if not ds.temporal_features:
        return chosen
    return sorted(chosen, key=lambda name: int(name.split("_")[-1]))


# --------------------------------------------------------------------------------------------
# Tree growth                                             src/random_forest_mc/model.py:222-293
# --------------------------------------------------------------------------------------------
def make_leaf(ds: OracleDataset, rows: np.ndarray, depth: int) -> dict:
    """model.py:222-229."""
    y = ds.columns[ds.target_col][rows]
    labels, counts = np.unique(y, return_counts=True)
    n = len(y)
    return {"leaf": {lab: cnt / n for lab, cnt in zip(labels, counts)}, "depth": f"{depth}#"}


def split_node(ds: OracleDataset, feat: str, rows: np.ndarray):
    """model.py:232-262.  Returns (rows for '>=', rows for '<', split value)."""
    v = ds.columns[feat][rows]
    if len(rows) <= 2:  # :258-262
        order = np.argsort(v)
        ranked = rows[order]
        return ranked[1:], ranked[:1], v[order][1]
    if feat in ds.numeric_cols:  # :236-245
        cut = float(np.quantile(v, 0.5))
        take = v >= cut
        hi, lo = rows[take], rows[~take]
        if len(hi) == 0 or len(lo) == 0:
            take = v > cut
            hi, lo = rows[take], rows[~take]
        return hi, lo, cut
    vals, counts = np.unique(v, return_counts=True)  # :249-254
    cut = vals[np.argmax(counts)]
    take = v == cut
    return rows[take], rows[~take], cut


def grow(ds: OracleDataset, rows: np.ndarray, feats: Sequence[str]) -> dict:
    """model.py:264-293, restated with an explicit work list instead of recursion.

    ``off`` is the rotation offset into the candidate's fixed feature list: the reference rotates
    a list copy per attempt (:276) and hands both children a copy of the rotated list (:287-288)."""
    nf = len(feats)
    y_all = ds.columns[ds.target_col]
    top: Dict[str, Any] = {}
    todo = [(rows, 0, 1, top, "root")]
    while todo:
        cur, off, depth, slot, key = todo.pop()
        if depth >= ds.max_depth or len(np.unique(y_all[cur])) == 1:  # :268
            slot[key] = make_leaf(ds, cur, depth)
            continue
        found = None
        for k in range(nf):  # :273-280 -- at most one full rotation
            feat = feats[(off + k) % nf]
            hi, lo, cut = split_node(ds, feat, cur)
            if len(hi) >= ds.min_samples_split and len(lo) >= ds.min_samples_split:
                found = (feat, hi, lo, cut, (off + k + 1) % nf)
                break
        if found is None:
            slot[key] = make_leaf(ds, cur, depth)
            continue
        feat, hi, lo, cut, nxt = found
        body = {"feat_type": ds.type_of_cols[feat], "split_val": cut, ">=": None, "<": None}  # key order :283-290
        slot[key] = {feat: {"split": body}}
        todo.append((lo, nxt, depth + 1, body, "<"))
        todo.append((hi, nxt, depth + 1, body, ">="))
    return top["root"]


# --------------------------------------------------------------------------------------------
# Traversal                                                src/random_forest_mc/tree.py:155-212
# --------------------------------------------------------------------------------------------
def walk(tree: dict, row) -> Tuple[dict, bool]:
    """tree.py:155-183 + :192-197.  Returns (leaf dict, saw_absent_column).

    When a split column is absent from the row the reference evaluates both subtrees (:171-175)
    but ``popLeafs`` returns from inside its ``for`` (:196-197), so only the first leaf -- the one
    reached by taking '>=' at every absent column -- is ever used."""
    is_series = isinstance(row, pd.Series)
    absent = False
    node = tree
    while True:
        name = next(iter(node))
        if name == "leaf":
            return node["leaf"], absent
        s = node[name]["split"]
        present = (name in row.index) if is_series else (name in row)
        if not present:
            absent = True
            node = s[">="]
            continue
        x = row[name]
        if x == s["split_val"] or (s["feat_type"] == "numeric" and x > s["split_val"]):  # :177-180
            node = s[">="]
        else:
            node = s["<"]


def tree_output(tree: dict, class_vals: Sequence[Any], row) -> dict:
    """tree.py:185-212: a plain leaf, or the renormalised single leaf on the absent-column path."""
    leaf, absent = walk(tree, row)
    if not absent:
        return leaf
    out = {c: 0 for c in class_vals}  # :201
    for c, p in leaf.items():
        out[c] += p
    for c in class_vals:
        out[c] /= 1  # :205-206 with len(leafes) == 1
    total = fsum(out.values())  # :208
    for c in class_vals:
        out[c] /= total
    return out


def top_class(dist: dict):
    """forest.py:200-202: stable descending sort -> first key among equal maxima."""
    return sorted(dist.items(), key=lambda kv: kv[1], reverse=True)[0][0]


# --------------------------------------------------------------------------------------------
# Candidate scoring + Monte-Carlo selection             src/random_forest_mc/model.py:296-348
# --------------------------------------------------------------------------------------------
def score_candidate(ds: OracleDataset, tree: dict, rows: np.ndarray):
    """model.py:296-314: accuracy of one candidate on its hold-out rows (np.float64)."""
    truth = ds.columns[ds.target_col][rows]
    guess = []
    for i in rows:
        row = {c: ds.columns[c][i] for c in ds.feature_cols}  # :309 all feature columns, numpy scalars
        guess.append(top_class(tree_output(tree, ds.class_vals, row)))
    return np.mean(truth == guess)


def select_tree(
    ds: OracleDataset,
    th_start: float = 1.0,
    delta_th: float = 0.1,
    max_discard_trees: int = 10,
    get_best_tree: bool = True,
    trace: Optional[list] = None,
):
    """model.py:317-348.  Returns (tree dict, survived_score, candidates grown).

    ``trace`` (if given) receives one record per candidate: rows, features, tree, score."""
    train_rows, val_rows = draw_rows(ds)
    threshold = th_start
    failures = 0
    best_tree, best_score = None, 0.0
    grown = 0
    while True:
        feats = draw_features(ds)
        cand = grow(ds, train_rows, list(feats))
        s = score_candidate(ds, cand, val_rows)
        grown += 1
        if trace is not None:
            trace.append({"idx_T": train_rows, "idx_V": val_rows, "F": list(feats), "tree": cand, "th_val": s})
        if s < threshold:  # :328
            if get_best_tree and best_score < s:  # :329-332 strict: first best wins
                best_tree, best_score = cand, s
        else:
            best_score = s  # :334
            break
        failures += 1
        if failures >= max_discard_trees:  # :336
            if get_best_tree:
                cand = best_tree  # may be None when every score was 0.0 (the reference then raises)
                break
            threshold -= delta_th  # :341
    return cand, best_score, grown


def fit(ds: OracleDataset, n_trees: int, trace: Optional[list] = None, **policy):
    """model.py:350-370.  Returns (list of tree dicts, list of survived scores, candidates grown)."""
    trees, scores, total = [], [], 0
    for _ in range(n_trees):
        t, s, g = select_tree(ds, trace=trace, **policy)
        trees.append(t)
        scores.append(s)
        total += g
    return trees, scores, total


# --------------------------------------------------------------------------------------------
# Voting                                                  src/random_forest_mc/forest.py:204-258
# --------------------------------------------------------------------------------------------
def forest_row(trees, scores, class_vals, row, soft: bool, weighted: bool) -> dict:
    """forest.py:204-233: the four voting modes for one row, fp64, sequential in tree order."""
    if soft:
        acc = defaultdict(float)
        if weighted:
            for t, s in zip(trees, scores):
                for c, p in tree_output(t, class_vals, row).items():
                    acc[c] += p * s  # :211 multiply, round, add, round
            denom = fsum(scores)
            return {c: acc[c] / denom for c in class_vals}
        outs = [tree_output(t, class_vals, row) for t in trees]
        for o in outs:
            for c, p in o.items():
                acc[c] += p
        return {c: acc[c] / len(outs) for c in class_vals}
    if weighted:
        acc = defaultdict(float)
        for t, s in zip(trees, scores):
            acc[top_class(tree_output(t, class_vals, row))] += s
        denom = fsum(scores)
        return {c: acc[c] / denom for c in class_vals}
    votes = [top_class(tree_output(t, class_vals, row)) for t in trees]
    return {c: votes.count(c) / len(votes) for c in class_vals}


def forest_probs(trees, scores, class_vals, frame: pd.DataFrame, soft: bool, weighted: bool) -> List[dict]:
    """forest.py:257-258."""
    return [forest_row(trees, scores, class_vals, row, soft, weighted) for _, row in frame.iterrows()]


def forest_labels(trees, scores, class_vals, frame: pd.DataFrame, soft: bool, weighted: bool) -> list:
    """forest.py:235-236."""
    return [top_class(forest_row(trees, scores, class_vals, row, soft, weighted)) for _, row in frame.iterrows()]
