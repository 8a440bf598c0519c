"""Lossless rank-code encoding of a training frame (the ``process_dataset`` step of the engine).

Reference boundary: ``RandomForestMC.process_dataset`` (src/random_forest_mc/model.py:162-195)
stores ``dataset_numpy = {col: ndarray}``; every later step of the hot path only ever *compares*
feature values (``>=``, ``>``, ``==``; model.py:238-252, tree.py:177-180) or takes order statistics
of them (np.quantile, np.unique+argmax, np.argsort; model.py:237,249-250,260).  Dense rank codes
preserve exactly those relations, so the GPU works on small integers:

* per column ``uniq = sorted distinct values``; ``code = rank + 1``; code 0 is the missing/unseen
  sentinel (never present in training rows because the reference drops NaN rows, model.py:181);
* numeric columns keep a float64 copy of ``uniq`` (exact for float32 and for |int| < 2**52) plus an
  arithmetic tag, because the split value of a >2-row node is *computed* (numpy's lerp in the
  column's own dtype) and has to be mapped back to a code threshold by a search in that table;
* categorical codes follow ``np.unique`` order so "lowest code" is the reference's tie-break for
  the mode (model.py:249-250);
* labels are coded by position in the lexicographically sorted class list (the key order of every
  leaf dict, model.py:224-227); ``sorted_to_classvals`` maps back to ``class_vals`` order.

HBM layout (one matrix for both kernels): row-major ``codes[n_rows][stride]`` in the narrowest
unsigned type that holds every column's codes, with the row pitch padded to an ODD number of 32-bit
words.  A bootstrap row is one contiguous, coalesced read for the grower's gather, and a tile of
consecutive rows copied verbatim into shared memory is bank-conflict-free for the traversal kernel
(thread r reads word r*pitch + f/2 -> bank (r + const) mod 32).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import pandas as pd

KIND_NUMERIC = 0
KIND_CATEGORICAL = 1

TAG_F64 = 0  # interpolate in float64
TAG_F32 = 1  # interpolate in float32 (numpy casts q to the array dtype)
TAG_INT = 2  # difference wraps in the integer dtype, then float64


def pack_tag(tag: int, bits: int = 0, signed: bool = False) -> int:
    return int(tag) | (int(bits) << 8) | (int(bool(signed)) << 16)


@dataclass
class ColumnCode:
    name: str
    kind: int
    uniq: np.ndarray  # sorted distinct values, native dtype (object array of str for categoricals)
    table: Optional[np.ndarray]  # float64 copy of uniq for numeric columns
    tag: int = 0  # packed: tag | bits<<8 | signed<<16

    @property
    def n_unique(self) -> int:
        return len(self.uniq)


@dataclass
class EncodedDataset:
    feature_cols: List[str]
    columns: List[ColumnCode]
    codes: np.ndarray  # [n_rows, stride] uint8/16/32, row-major, pitch = odd number of 32-bit words
    width: int  # bytes per code
    stride: int  # elements per row
    labels: np.ndarray  # uint8 [n_rows]: index into class_sorted
    class_sorted: List  # labels in np.unique order (leaf-dict key order)
    class_vals: List  # first-appearance order (model.py:184)
    sorted_to_classvals: np.ndarray  # int32 [C]
    class_rows: List[np.ndarray] = field(default_factory=list)  # per class_vals entry: ascending row ids

    @property
    def n_rows(self) -> int:
        return self.codes.shape[0]

    @property
    def n_features(self) -> int:
        return len(self.feature_cols)

    @property
    def n_classes(self) -> int:
        return len(self.class_sorted)


def odd_word_stride(n_features: int, width: int) -> int:
    """Row pitch in elements such that pitch*width is an odd multiple of 4 bytes."""
    words = (n_features * width + 3) // 4
    if words % 2 == 0:
        words += 1
    return words * 4 // width


def _rank(values: np.ndarray):
    """(sorted distinct values, 0-based rank of every element).  Same order as np.unique."""
    # factorize is O(n) hashing; sorting only the distinct values keeps np.unique's order
    # (object arrays sort with Python comparisons in both, so "10" < "2" as in the reference)
    inv0, firsts = pd.factorize(values, sort=False)
    firsts = np.asarray(firsts, dtype=values.dtype)
    order = np.argsort(firsts, kind="stable")
    rank_of = np.empty(len(firsts), dtype=np.int64)
    rank_of[order] = np.arange(len(firsts))
    return firsts[order], rank_of[inv0]


def encode_column(name: str, values: np.ndarray, numeric: bool):
    uniq, rank = _rank(values)
    if numeric and values.dtype.kind == "f":
        # -0.0 and 0.0 compare equal and share one code; which of the two the reference prints as a
        # split value depends on numpy's introselect element order, so the engine always emits +0.0
        uniq = uniq + uniq.dtype.type(0)
    if not numeric:
        return ColumnCode(name, KIND_CATEGORICAL, uniq, None, 0), rank
    dt = values.dtype
    if dt == np.float64:
        tag = pack_tag(TAG_F64)
    elif dt == np.float32:
        tag = pack_tag(TAG_F32)
    elif dt.kind in "iu":
        if len(uniq) and max(abs(int(uniq[0])), abs(int(uniq[-1]))) >= 2**52:
            raise NotImplementedError(
                f"column {name!r}: integer magnitudes >= 2**52 are not representable in the float64 value table"
            )
        tag = pack_tag(TAG_INT, dt.itemsize * 8, dt.kind == "i")
    else:
        raise NotImplementedError(f"column {name!r}: numeric dtype {dt} is not supported by the B200 engine")
    return ColumnCode(name, KIND_NUMERIC, uniq, uniq.astype(np.float64), tag), rank


def encode_dataset(
    dataset_numpy: Dict[str, np.ndarray],
    feature_cols: Sequence[str],
    numeric_cols: Sequence[str],
    target_col: str,
    class_vals: Sequence,
) -> EncodedDataset:
    """Encode the (already NaN-free) column dict produced by ``process_dataset``."""
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
    y = dataset_numpy[target_col]
    class_sorted, yrank = _rank(y)
    if len(class_sorted) > 255:
        raise NotImplementedError("more than 255 classes")
    class_sorted = list(class_sorted)
    s2c = np.array([list(class_vals).index(c) for c in class_sorted], dtype=np.int32)
    labels = yrank.astype(np.uint8)
    rows_by_sorted = [np.flatnonzero(labels == k) for k in range(len(class_sorted))]
    class_rows = [rows_by_sorted[class_sorted.index(c)] for c in class_vals]
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


def median_value(va: float, vb: float, odd: bool, tag: int) -> float:
    """numpy's linear-interpolation median of two adjacent order statistics, in the column's dtype.

    Restates numpy ``_lerp`` (numpy/lib/_function_base_impl.py:4657-4678 in numpy 2.3.5) for q=0.5 as
    called from model.py:237: gamma is 0 for odd n (``a + (b-a)*0``) and 0.5 for even n
    (``b - (b-a)*0.5``); float32 columns are interpolated in float32; for integer columns ``b-a``
    wraps in the integer dtype before the float64 multiply.  The CUDA grower computes the same."""
    kind = tag & 0xFF
    if kind == TAG_F32:
        a, b = np.float32(va), np.float32(vb)
        with np.errstate(all="ignore"):
            d = np.float32(b - a)
            r = np.float32(a + np.float32(d * np.float32(0.0))) if odd else np.float32(b - np.float32(d * np.float32(0.5)))
        return float(r)
    a, b = np.float64(va), np.float64(vb)
    with np.errstate(all="ignore"):
        d = np.float64(b - a)
        if kind == TAG_INT and (tag >> 16) & 1:
            half = np.float64(2.0 ** (((tag >> 8) & 0xFF) - 1))
            if d >= half:
                d = np.float64(d - 2 * half)
        r = a + d * np.float64(0.0) if odd else b - d * np.float64(0.5)
    return float(r)
