"""TEST INFRASTRUCTURE -- numpy restatement of the rank coding that `rfmc_dataset_encode` does on the
GPU (process_dataset, src/random_forest_mc/model.py:162-195 -> the code matrix of DESIGN.md section 3).

Only tests/, __graft_entry__.smoke() and tests/fake_engine.py import this; the product encodes on
the device and has no host path for numeric columns.

Pinned: tests/test_code_model.py checks that these codes are lossless against the reference's
recorded `dataset_numpy` columns (values round-trip through the tables; growing and voting on the
codes reproduces every recorded reference tree and prediction of tests/golden/*.pkl.gz).

Per column: ``uniq = np.unique(values)`` (model.py:249 uses the same order for categoricals),
``code = rank + 1``, 0 = missing/unseen.  Numeric columns also carry ``uniq`` as float64 and the tag
of the dtype numpy's quantile interpolates in (model.py:237)."""

from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from random_forest_mc_b200.encode import (
    KIND_CATEGORICAL,
    KIND_NUMERIC,
    ColumnCode,
    EncodedDataset,
    check_int_range,
    encode_labels,
    numeric_tag,
    odd_word_stride,
)


def rank_numpy(values: np.ndarray):
    """(sorted distinct values, 0-based rank of every element) straight from np.unique."""
    uniq, inv = np.unique(values, return_inverse=True)
    return uniq, inv.reshape(-1).astype(np.int64)


def encode_column(name: str, values: np.ndarray, numeric: bool):
    uniq, rank = rank_numpy(values)
    if not numeric:
        return ColumnCode(name, KIND_CATEGORICAL, uniq, None, 0), rank
    if values.dtype.kind == "f":
        # -0.0 and 0.0 compare equal and share one code; the engine always emits +0.0
        uniq = uniq + uniq.dtype.type(0)
    tag = numeric_tag(name, values.dtype)
    if values.dtype.kind in "iu" and len(uniq):
        check_int_range(name, float(uniq[0]), float(uniq[-1]))
    return ColumnCode(name, KIND_NUMERIC, uniq, uniq.astype(np.float64), tag), rank


def encode_dataset(
    dataset_numpy: Dict[str, np.ndarray],
    feature_cols: Sequence[str],
    numeric_cols: Sequence[str],
    target_col: str,
    class_vals: Sequence,
) -> EncodedDataset:
    """Encode the (already NaN-free) column dict produced by ``process_dataset`` on the host."""
    numeric = set(numeric_cols)
    cols: List[ColumnCode] = []
    ranks = []
    for name in feature_cols:
        cc, rank = encode_column(name, dataset_numpy[name], name in numeric)
        cols.append(cc)
        ranks.append(rank)
    n = len(dataset_numpy[target_col])
    top = max([c.n_unique for c in cols] + [1])
    if top <= 0xFF:
        dtype, width = np.uint8, 1
    elif top <= 0xFFFF:
        dtype, width = np.uint16, 2
    else:
        dtype, width = np.uint32, 4
    stride = odd_word_stride(len(cols), width)
    codes = np.zeros((n, stride), dtype=dtype)
    for j, rank in enumerate(ranks):
        codes[:, j] = rank + 1
    class_sorted, labels, s2c, class_rows = encode_labels(dataset_numpy[target_col], class_vals)
    return EncodedDataset(
        feature_cols=list(feature_cols),
        columns=cols,
        codes=codes,
        width=width,
        stride=stride,
        labels=labels,
        class_sorted=class_sorted,
        class_vals=list(class_vals),
        sorted_to_classvals=s2c,
        class_rows=class_rows,
    )
