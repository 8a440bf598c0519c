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

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

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


class EncodedDataset:
    """Host-side description of the rank-coded training matrix.  The matrix itself lives in HBM
    (``engine.DeviceDataset``); ``codes`` is a host copy fetched on first use (tests, inspection)."""

    def __init__(
        self,
        feature_cols: List[str],
        columns: List[ColumnCode],
        codes: Optional[np.ndarray],  # [n_rows, stride] uint8/16/32, row-major, pitch = odd number of 32-bit words
        width: int,  # bytes per code
        stride: int,  # elements per row
        labels: np.ndarray,  # uint8 [n_rows]: index into class_sorted
        class_sorted: List,  # labels in np.unique order (leaf-dict key order)
        class_vals: List,  # first-appearance order (model.py:184)
        sorted_to_classvals: np.ndarray,  # int32 [C]
        class_rows: Optional[List[np.ndarray]] = None,  # per class_vals entry: ascending row ids
        fetch_codes: Optional[Callable[[], np.ndarray]] = None,
    ):
        self.feature_cols = feature_cols
        self.columns = columns
        self._codes = codes
        self._fetch_codes = fetch_codes
        self.width = width
        self.stride = stride
        self.labels = labels
        self.class_sorted = class_sorted
        self.class_vals = class_vals
        self.sorted_to_classvals = sorted_to_classvals
        self.class_rows = class_rows if class_rows is not None else []

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_fetch_codes"] = None  # a closure over a device handle: a copy keeps the description, not the matrix
        return state

    @property
    def codes(self) -> np.ndarray:
        if self._codes is None:
            if self._fetch_codes is None:
                raise RuntimeError("the code matrix of this dataset is not available on the host")
            self._codes = self._fetch_codes()
        return self._codes

    @property
    def n_rows(self) -> int:
        return len(self.labels)

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


class LazyColumns(dict):
    """``dataset_numpy`` (model.py:194): ``{col: ndarray}``.  Numeric columns are views of the frame and
    are stored right away; string columns (Arrow-backed in pandas 3) become object arrays only when
    someone reads them -- the engine ranks them inside Arrow and never needs the object arrays."""

    def __init__(self, frame: pd.DataFrame, eager: Sequence[str]):
        super().__init__()
        self._frame = frame
        self._order = list(frame.columns)
        for c in eager:
            dict.__setitem__(self, c, frame[c].to_numpy())

    def __missing__(self, key):
        if key not in self._order:
            raise KeyError(key)
        v = self._frame[key].to_numpy()
        dict.__setitem__(self, key, v)
        return v

    def column(self, key) -> pd.Series:
        return self._frame[key]

    def __contains__(self, key):
        return key in self._order

    def __iter__(self):
        return iter(self._order)

    def __len__(self):
        return len(self._order)

    def keys(self):
        return list(self._order)

    def values(self):
        return [self[k] for k in self._order]

    def items(self):
        return [(k, self[k]) for k in self._order]

    def get(self, key, default=None):
        return self[key] if key in self._order else default


def rank_series(col: pd.Series):
    """rank_values for a frame column: Series.factorize keeps Arrow-backed strings in Arrow
    (dictionary encode, GIL released); only the distinct values become Python objects."""
    inv0, firsts = col.factorize(use_na_sentinel=False)
    firsts = firsts.to_numpy()  # the dtype Series.to_numpy() would give (model.py:194): object for strings, bool for bool
    order = np.argsort(firsts, kind="stable")
    rank_of = np.empty(len(firsts), dtype=np.int64)
    rank_of[order] = np.arange(len(firsts))
    return firsts[order], rank_of[inv0]


def rank_values(values: np.ndarray):
    """(sorted distinct values, 0-based rank of every element).  Same order as np.unique.  Host code,
    used for categorical columns and the target: strings never go to the GPU."""
    # factorize is O(n) hashing; sorting only the distinct values keeps np.unique's order
    # (object arrays sort with Python comparisons in both, so "10" < "2" as in the reference)
    inv0, firsts = pd.factorize(values, sort=False)
    firsts = np.asarray(firsts, dtype=values.dtype)
    order = np.argsort(firsts, kind="stable")
    rank_of = np.empty(len(firsts), dtype=np.int64)
    rank_of[order] = np.arange(len(firsts))
    return firsts[order], rank_of[inv0]


def numeric_tag(name: str, dt: np.dtype) -> int:
    """Packed arithmetic tag of a numeric column: the dtype numpy's median interpolates in."""
    if dt == np.float64:
        return pack_tag(TAG_F64)
    if dt == np.float32:
        return pack_tag(TAG_F32)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        return pack_tag(TAG_INT, dt.itemsize * 8, dt.kind == "i")
    raise NotImplementedError(f"column {name!r}: numeric dtype {dt} is not supported by the B200 engine")


def check_int_range(name: str, lo: float, hi: float) -> None:
    if max(abs(lo), abs(hi)) >= 2**52:
        raise NotImplementedError(
            f"column {name!r}: integer magnitudes >= 2**52 are not representable in the float64 value table"
        )


def encode_labels(y, class_vals: Sequence):
    """Target column -> (class_sorted, labels uint8, sorted_to_classvals, class_rows).  Labels are
    coded in lexicographic order (the key order of every leaf dict, model.py:224-227); class_rows
    are the cached ``indices[target_vals == val]`` of model.py:205, in ``class_vals`` order."""
    class_sorted, yrank = rank_series(y) if isinstance(y, pd.Series) else rank_values(y)
    if len(class_sorted) > 255:
        raise NotImplementedError("more than 255 classes")
    class_sorted = list(class_sorted)
    s2c = np.array([list(class_vals).index(c) for c in class_sorted], dtype=np.int32)
    labels = yrank.astype(np.uint8)
    rows_by_sorted = [np.flatnonzero(labels == k) for k in range(len(class_sorted))]
    class_rows = [rows_by_sorted[class_sorted.index(c)] for c in class_vals]
    return class_sorted, labels, s2c, class_rows


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
