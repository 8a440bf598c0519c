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
    """Materialise the nested dict the reference would have built for this tree."""
    nodes, counts = tree.nodes, tree.counts
    feat, thr, child, flags, sval = (nodes[k] for k in ("feat", "thr", "child", "flags", "sval"))
    labels = enc.class_sorted
    holder = {}
    todo = [(0, 1, holder, "root")]
    while todo:
        i, depth, slot, key = todo.pop()
        f = int(feat[i])
        if f < 0:
            row = counts[i]
            total = int(row.sum())
            leaf = {labels[c]: np.int64(row[c]) / total for c in np.flatnonzero(row)}
            slot[key] = {"leaf": leaf, "depth": f"{depth}#"}
            continue
        col = enc.columns[f]
        fl = int(flags[i])
        if fl & FLAG_CATEGORICAL or fl & FLAG_TWO_ROW:
            value = col.uniq[int(thr[i]) - 1]
        else:
            value = float(sval[i])
        kind = "categorical" if col.kind == KIND_CATEGORICAL else "numeric"
        if type_of_cols is not None:
            kind = type_of_cols[col.name]
        body = {"feat_type": kind, "split_val": value, ">=": None, "<": None}
        slot[key] = {col.name: {"split": body}}
        c = int(child[i])
        todo.append((c + 1, depth + 1, body, "<"))
        todo.append((c, depth + 1, body, ">="))
    return holder["root"]


def used_feature_ids(tree: TreeSoA) -> List[int]:
    f = tree.nodes["feat"]
    return sorted(set(int(x) for x in np.unique(f[f >= 0])))
