"""ctypes binding of librfmc_b200.so (include/rfmc_b200.h) -- the only door to the GPU.

There is no CPU fallback: if the shared library is missing or no CUDA device is visible, every
compute entry point raises ``EngineUnavailable``.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from .encode import (
    KIND_CATEGORICAL,
    KIND_NUMERIC,
    ColumnCode,
    EncodedDataset,
    check_int_range,
    encode_labels,
    numeric_tag,
    rank_series,
    rank_values,
)
from .flat import NODE_DTYPE, TreeSoA

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librfmc_b200.so")

PNODE_DTYPE = np.dtype([("feat_flags", "<u4"), ("thr", "<u4"), ("child", "<u4"), ("reserved", "<u4")])
PNODE_LEAF = 0x80000000
PNODE_CATEGORICAL = 0x40000000
MAX_CLASSES = 64


class EngineUnavailable(RuntimeError):
    pass


class EngineError(RuntimeError):
    pass


_lib = None

_p = C.c_void_p
_SIGNATURES = {
    "rfmc_abi_version": (C.c_int, []),
    "rfmc_last_error": (C.c_char_p, []),
    "rfmc_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "rfmc_cache_trim": (C.c_int, [C.c_int, C.POINTER(C.c_int64)]),
    "rfmc_stream_create": (C.c_int, [C.c_int, C.POINTER(_p)]),
    "rfmc_stream_destroy": (C.c_int, [C.c_int, _p]),
    "rfmc_dataset_create": (C.c_int, [C.c_int, _p, C.c_int, C.c_int64, C.c_int32, C.c_int64, _p, C.c_int32, _p, _p, _p, _p, _p, C.POINTER(_p)]),
    "rfmc_dataset_destroy": (C.c_int, [_p]),
    "rfmc_dataset_encode": (C.c_int, [C.c_int, C.c_int64, C.c_int32, _p, _p, _p, _p, _p, C.c_int32, C.POINTER(_p)]),
    "rfmc_dataset_describe": (C.c_int, [_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), _p, _p]),
    "rfmc_dataset_tables": (C.c_int, [_p, _p]),
    "rfmc_dataset_codes": (C.c_int, [_p, _p]),
    "rfmc_batch_create": (C.c_int, [_p, C.c_int32, C.c_int32, C.c_int32, _p, _p, C.c_int32, _p, _p, _p, C.c_int32, C.c_int32, _p, C.POINTER(_p)]),
    "rfmc_batch_run": (C.c_int, [_p, _p]),
    "rfmc_batch_results": (C.c_int, [_p, _p, _p]),
    "rfmc_batch_pack": (C.c_int, [_p, _p, C.c_int32, _p, _p, C.c_int, _p]),
    "rfmc_batch_destroy": (C.c_int, [_p]),
    "rfmc_batch_launches": (C.c_int, [_p]),
    "rfmc_batch_set_timing": (C.c_int, [_p, C.c_int]),
    "rfmc_batch_kernel_ms": (C.c_int, [_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "rfmc_forest_create": (C.c_int, [C.c_int, _p, C.c_int64, _p, C.c_int32, C.c_int64, _p, _p, _p, _p, C.c_int32, _p, C.c_double, C.POINTER(_p)]),
    "rfmc_forest_destroy": (C.c_int, [_p]),
    "rfmc_forest_predict": (C.c_int, [_p, _p, C.c_int, C.c_int64, C.c_int64, C.c_int32, _p, C.c_int, C.c_int, _p, _p, C.c_int, _p]),
    "rfmc_forest_leaves": (C.c_int, [_p, _p, C.c_int, C.c_int64, C.c_int64, C.c_int32, _p, _p, _p]),
    "rfmc_legacy_permutation_prefix": (C.c_int, [_p, C.POINTER(C.c_int32), C.c_int64, C.c_int64, _p]),
    "rfmc_pyrandom_sample_plan": (C.c_int, [_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, _p, _p]),
    "rfmc_forest_set_codebook": (C.c_int, [_p, C.c_int32, _p, _p]),
    "rfmc_forest_layout": (C.c_int, [_p, _p]),
    "rfmc_forest_predict_frame": (C.c_int, [_p, C.c_int64, C.c_int32, C.c_int64, C.c_int, C.c_int32, _p, _p, _p, _p,
                                            C.c_int, C.c_int, _p, _p]),
    "rfmc_legacy_draw_plan": (C.c_int, [_p, C.POINTER(C.c_int32), C.c_int32, C.c_int32, _p, _p, C.c_int64, C.c_int64,
                                        C.c_int32, C.c_int64, C.c_int64, _p, _p, _p, _p, _p, C.c_int32]),
    "rfmc_mt_jump_poly": (C.c_int, [C.c_int64, _p]),
    "rfmc_mt_jump_window_host": (C.c_int, [_p, C.c_int64, _p]),
    "rfmc_dataset_set_class_rows": (C.c_int, [_p, C.c_int32, _p, _p]),
    "rfmc_draw_plan_device": (C.c_int, [_p, _p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64, _p,
                                        C.POINTER(_p)]),
    "rfmc_drawplan_finish": (C.c_int, [_p, _p, _p, C.POINTER(C.c_int32)]),
    "rfmc_drawplan_state_at": (C.c_int, [_p, C.c_int32, _p, C.POINTER(C.c_int32)]),
    "rfmc_drawplan_indices": (C.c_int, [_p, C.c_int32, C.c_int32, _p, _p]),
    "rfmc_drawplan_info": (C.c_int, [_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32)]),
    "rfmc_drawplan_timing": (C.c_int, [C.c_int]),
    "rfmc_drawplan_kernel_ms": (C.c_int, [_p, _p]),
    "rfmc_drawplan_destroy": (C.c_int, [_p]),
    "rfmc_batch_create_from_plan": (C.c_int, [_p, _p, C.c_int32, _p, _p, _p, C.c_int32, C.c_int32, _p, C.POINTER(_p)]),
    "rfmc_allgather_survivors": (C.c_int, [_p, C.c_int, _p, _p, C.c_int64, C.c_int32, _p, _p, _p, C.c_int32, C.c_int32, _p, C.POINTER(_p)]),
    "rfmc_gathered_describe": (C.c_int, [_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int64), _p, _p]),
    "rfmc_gathered_meta": (C.c_int, [_p, C.c_int32, _p, _p, _p]),
    "rfmc_gathered_fetch": (C.c_int, [_p, C.c_int32, _p, _p]),
    "rfmc_gathered_checksum": (C.c_int, [_p, C.POINTER(C.c_uint64)]),
    "rfmc_gathered_destroy": (C.c_int, [_p]),
}
DRAW_MAX_PER_CLASS = 4096  # RFMC_DRAW_MAX_PER_CLASS
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load_library():
    """Load the C-ABI library (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineUnavailable(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  The engine has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _check(rc: int):
    if rc != 0:
        raise EngineError(load_library().rfmc_last_error().decode("utf-8", "replace"))


def cache_trim(device: int = 0) -> int:
    """Give the library's cached device blocks and pinned staging buffers back to the driver; bytes freed."""
    n = C.c_int64(0)
    _check(load_library().rfmc_cache_trim(int(device), C.byref(n)))
    return int(n.value)


def device_count() -> int:
    n = C.c_int(0)
    try:
        rc = load_library().rfmc_device_count(C.byref(n))
    except EngineUnavailable:
        raise
    if rc != 0:
        return 0
    return n.value


def require_gpu():
    if device_count() < 1:
        raise EngineUnavailable("no CUDA device visible: the B200 engine has no CPU fallback")


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_p)


def _stream_ptr(stream) -> Optional[int]:
    if stream is None:
        return None
    if isinstance(stream, int):
        return stream or None
    return getattr(stream, "cuda_stream", None) or None


def stream_create(device: int = 0) -> int:
    h = _p()
    _check(load_library().rfmc_stream_create(device, C.byref(h)))
    return h.value


def stream_destroy(device: int, stream: int) -> None:
    if stream:
        _check(load_library().rfmc_stream_destroy(device, _p(stream)))


_STREAM_POOL = {}
_STREAM_LOCK = __import__("threading").Lock()


def stream_acquire(device: int = 0) -> int:
    """A CUDA stream for one host thread of a multi-stream fit.  Streams are kept and reused:
    creating and (above all) destroying one while other streams are busy costs milliseconds."""
    with _STREAM_LOCK:
        free = _STREAM_POOL.setdefault(device, [])
        if free:
            return free.pop()
    return stream_create(device)


def stream_release(device: int, stream: int) -> None:
    if stream:
        with _STREAM_LOCK:
            _STREAM_POOL.setdefault(device, []).append(stream)


def legacy_choice_indices(key: np.ndarray, pos: int, n: int, size: int):
    """permutation(n)[:size] on the legacy MT19937 stream (key is advanced in place).  Host code."""
    out = np.empty(size, dtype=np.int64)
    p = C.c_int32(pos)
    rc = load_library().rfmc_legacy_permutation_prefix(_ptr(key), C.byref(p), n, size, _ptr(out))
    if rc != 0:
        raise EngineError("rfmc_legacy_permutation_prefix: bad arguments")
    return out, p.value


_FLAT_ROWS = {}


def _flat_class_rows(class_rows):
    """(rows back to back, offsets) of a dataset's per-class row lists; built once per list object
    (a fit calls the draw plan hundreds of times, from several host threads)."""
    hit = _FLAT_ROWS.get(id(class_rows))
    if hit is not None and hit[0] is class_rows:
        return hit[1], hit[2]
    rows = np.ascontiguousarray(np.concatenate(class_rows), dtype=np.int64)
    off = np.zeros(len(class_rows) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in class_rows])
    if len(_FLAT_ROWS) > 8:
        _FLAT_ROWS.clear()
    _FLAT_ROWS[id(class_rows)] = (class_rows, rows, off)
    return rows, off


def legacy_draw_plan(key: np.ndarray, pos: int, n_slots: int, class_rows: Sequence[np.ndarray], per_class: int,
                     n_train_pclass: int, n_cands: int, low: int, high: int, n_threads: int = 4):
    """The numpy-stream draws of ``n_slots`` tree slots (rows per class, then ``n_cands`` feature
    counts per slot) in one host call.  Returns (idx_train, idx_val, feat_counts, key_snap,
    pos_snap, new_pos); ``key`` is advanced in place."""
    C_ = len(class_rows)
    rows, off = _flat_class_rows(class_rows)
    n_tr = min(per_class, n_train_pclass)
    idx_train = np.empty((n_slots, C_ * n_tr), dtype=np.int32)
    idx_val = np.empty((n_slots, C_ * (per_class - n_tr)), dtype=np.int32)
    counts = np.empty((n_slots, n_cands), dtype=np.int32)
    key_snap = np.empty((n_slots, 624), dtype=np.uint32)
    pos_snap = np.empty(n_slots, dtype=np.int32)
    p = C.c_int32(pos)
    rc = load_library().rfmc_legacy_draw_plan(
        _ptr(key), C.byref(p), n_slots, C_, _ptr(rows), _ptr(off), per_class, n_train_pclass, n_cands, low, high,
        _ptr(idx_train), _ptr(idx_val), _ptr(counts), _ptr(key_snap), _ptr(pos_snap), n_threads,
    )
    if rc != 0:
        raise EngineError("rfmc_legacy_draw_plan: bad arguments")
    return idx_train, idx_val, counts, key_snap, pos_snap, p.value


def pyrandom_sample_plan(rng, n_pop: int, ks: np.ndarray):
    """``rng.sample(range(n_pop), k)`` for every k of ``ks`` in a row, on ``rng``'s own stream
    (``rng`` = the ``random`` module or a ``random.Random``).  Returns the picked indices back to
    back; the generator is left exactly where CPython would have left it."""
    version, internal, gauss = rng.getstate()
    key = np.array(internal[:624], dtype=np.uint32)
    p = C.c_int32(int(internal[624]))
    ks = np.ascontiguousarray(ks, dtype=np.int32).reshape(-1)
    out = np.empty(int(ks.sum()), dtype=np.int32)
    rc = load_library().rfmc_pyrandom_sample_plan(_ptr(key), C.byref(p), n_pop, len(ks), _ptr(ks), _ptr(out))
    if rc != 0:
        raise EngineError("rfmc_pyrandom_sample_plan: bad arguments")
    rng.setstate((version, tuple(key.tolist()) + (p.value,), gauss))
    return out


_CLASS_ROWS_LOCK = __import__("threading").Lock()


class DeviceDataset:
    """rfmc_dataset handle: the encoded training matrix resident in HBM (process_dataset)."""

    def __init__(self, enc: EncodedDataset, device: int = 0):
        require_gpu()
        self.enc = enc
        self.device = device
        nfeat = enc.n_features
        kind = np.array([c.kind for c in enc.columns], dtype=np.uint8)
        tag = np.array([c.tag for c in enc.columns], dtype=np.int32)
        nuniq = np.array([c.n_unique for c in enc.columns], dtype=np.int32)
        tabs = [c.table if c.kind != KIND_CATEGORICAL else np.zeros(0) for c in enc.columns]
        off = np.zeros(nfeat + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(t) for t in tabs])
        tables = np.ascontiguousarray(np.concatenate(tabs) if tabs else np.zeros(0), dtype=np.float64)
        if len(tables) == 0:
            tables = np.zeros(1, dtype=np.float64)
        codes = np.ascontiguousarray(enc.codes)
        labels = np.ascontiguousarray(enc.labels, dtype=np.uint8)
        h = _p()
        _check(
            load_library().rfmc_dataset_create(
                device, _ptr(codes), enc.width, codes.shape[0], nfeat, enc.stride, _ptr(labels), enc.n_classes,
                _ptr(kind), _ptr(tag), _ptr(nuniq), _ptr(off), _ptr(tables), C.byref(h),
            )
        )
        self._h = h

    @classmethod
    def from_handle(cls, handle, device: int = 0) -> "DeviceDataset":
        """Wrap a dataset the library built itself (rfmc_dataset_encode)."""
        self = cls.__new__(cls)
        self.enc = None
        self.device = device
        self._h = handle
        return self

    def set_class_rows(self, class_rows: Sequence[np.ndarray]) -> None:
        """Per-class row lists (class_vals order) for the device draws; uploaded once."""
        if getattr(self, "_class_rows_id", None) is class_rows:
            return
        with _CLASS_ROWS_LOCK:  # the host streams of a fit reach this together: one upload, before any plan uses it
            if getattr(self, "_class_rows_id", None) is class_rows:
                return
            rows, off = _flat_class_rows(class_rows)
            _check(load_library().rfmc_dataset_set_class_rows(self._h, len(class_rows), _ptr(rows), _ptr(off)))
            self._class_rows_id = class_rows

    def close(self):
        if getattr(self, "_h", None):
            load_library().rfmc_dataset_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DrawPlan:
    """rfmc_drawplan handle: the numpy-stream draws of a plan of tree slots, walked on the GPU
    (split_train_val / sampleFeats, model.py:198-219).  Row ids stay in HBM for the grower."""

    def __init__(self, ds: DeviceDataset, class_rows: Sequence[np.ndarray], key: np.ndarray, pos: int, n_slots: int,
                 per_class: int, n_train_pclass: int, n_cands: int, low: int, high: int, stream=None):
        self.ds = ds
        ds.set_class_rows(class_rows)
        self.n_slots, self.n_cands = int(n_slots), int(n_cands)
        self.n_classes = len(class_rows)
        self.n_tr = min(int(per_class), int(n_train_pclass))
        self.n_va = int(per_class) - self.n_tr
        self._stream = _stream_ptr(stream)
        key = np.ascontiguousarray(key, dtype=np.uint32)
        h = _p()
        _check(load_library().rfmc_draw_plan_device(ds._h, _ptr(key), int(pos), self.n_slots, int(per_class), int(n_train_pclass),
                                                     self.n_cands, int(low), int(high), self._stream, C.byref(h)))
        self._h = h

    def finish(self):
        """Wait for the plan.  Returns (feat_counts [n_slots, n_cands], key_end, pos_end)."""
        counts = np.zeros((self.n_slots, self.n_cands), dtype=np.int32)
        key = np.empty(624, dtype=np.uint32)
        pos = C.c_int32(0)
        _check(load_library().rfmc_drawplan_finish(self._h, _ptr(counts), _ptr(key), C.byref(pos)))
        return counts, key, pos.value

    def state_at(self, slot: int):
        """numpy's stream state (key, pos) before slot ``slot`` (``n_slots``: after the plan)."""
        key = np.empty(624, dtype=np.uint32)
        pos = C.c_int32(0)
        _check(load_library().rfmc_drawplan_state_at(self._h, int(slot), _ptr(key), C.byref(pos)))
        return key, pos.value

    def indices(self, slot0: int = 0, n_slots: Optional[int] = None):
        n = self.n_slots - slot0 if n_slots is None else int(n_slots)
        T = np.empty((n, self.n_classes * self.n_tr), dtype=np.int32)
        V = np.empty((n, self.n_classes * self.n_va), dtype=np.int32)
        _check(load_library().rfmc_drawplan_indices(self._h, int(slot0), n, _ptr(T), _ptr(V)))
        return T, V

    def info(self):
        g, c, p, f, l = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int32(0), C.c_int32(0)
        _check(load_library().rfmc_drawplan_info(self._h, C.byref(g), C.byref(c), C.byref(p), C.byref(f), C.byref(l)))
        return dict(words_generated=g.value, words_consumed=c.value, words_pending=p.value, chunk_fallbacks=f.value, launches=l.value)

    def kernel_ms(self):
        """(fill, classify + chain rounds, final pass, resolve) in ms; needs engine.drawplan_timing(True) before creation."""
        ms = (C.c_float * 4)()
        _check(load_library().rfmc_drawplan_kernel_ms(self._h, C.cast(ms, _p)))
        return tuple(float(x) for x in ms)

    def close(self):
        if getattr(self, "_h", None):
            load_library().rfmc_drawplan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def drawplan_timing(enabled: bool) -> None:
    _check(load_library().rfmc_drawplan_timing(int(bool(enabled))))


def mt_jump_window_host(key: np.ndarray, offset: int) -> np.ndarray:
    """Host evaluation of the MT19937 jump: the untempered window x[offset .. offset+623]."""
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(624, dtype=np.uint32)
    if load_library().rfmc_mt_jump_window_host(_ptr(key), int(offset), _ptr(out)) != 0:
        raise EngineError("rfmc_mt_jump_window_host: bad arguments")
    return out


# numpy dtype -> RFMC_DT_* (include/rfmc_b200.h)
_RAW_DTYPES = {"f8": 0, "f4": 1, "i8": 2, "i4": 3, "i2": 4, "i1": 5, "u1": 6, "u2": 7, "u4": 8, "u8": 9}
DT_CODE = 10


def encode_dataset_device(dataset_numpy, feature_cols, numeric_cols, target_col, class_vals, device: int = 0):
    """``process_dataset``'s encode step on the GPU (rfmc_dataset_encode): numeric columns are sent
    raw and rank-coded on the device; categorical columns and the target are ranked on the host
    (strings) and sent as codes.  Returns (EncodedDataset, DeviceDataset); the code matrix stays in
    HBM and is copied to the host only if someone reads ``EncodedDataset.codes``."""
    require_gpu()
    numeric = set(numeric_cols)
    series_of = getattr(dataset_numpy, "column", None)  # LazyColumns: rank strings inside Arrow
    n = len(series_of(target_col)) if series_of else len(dataset_numpy[target_col])
    nfeat = len(feature_cols)
    keep = []  # host buffers the pointers refer to
    ptrs = (C.c_void_p * nfeat)()
    dtypes = np.zeros(nfeat, dtype=np.int32)
    tags = np.zeros(nfeat, dtype=np.int32)
    cat_n = np.zeros(nfeat, dtype=np.int32)
    cat_uniq = {}
    def rank_categorical(name):
        uniq, rank = rank_series(series_of(name)) if series_of else rank_values(dataset_numpy[name])
        return uniq, (rank + 1).astype(np.int32)

    cat_names = [name for name in feature_cols if name not in numeric]
    if len(cat_names) > 1 and n >= 100_000:
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=min(8, len(cat_names))) as pool:
            ranked = dict(zip(cat_names, pool.map(rank_categorical, cat_names)))
    else:
        ranked = {name: rank_categorical(name) for name in cat_names}
    for j, name in enumerate(feature_cols):
        if name in numeric:
            v = dataset_numpy[name]
            tags[j] = numeric_tag(name, v.dtype)
            v = np.ascontiguousarray(v, dtype=v.dtype.newbyteorder("="))
            dtypes[j] = _RAW_DTYPES[v.dtype.str[1:]]
        else:
            cat_uniq[j], v = ranked[name]
            cat_n[j] = len(cat_uniq[j])
            dtypes[j] = DT_CODE
        keep.append(v)
        ptrs[j] = v.ctypes.data
    class_sorted, labels, s2c, class_rows = encode_labels(
        series_of(target_col) if series_of else dataset_numpy[target_col], class_vals
    )
    if len(class_sorted) > MAX_CLASSES:
        raise NotImplementedError(f"more than {MAX_CLASSES} classes")
    h = _p()
    _check(
        load_library().rfmc_dataset_encode(
            device, n, nfeat, C.cast(ptrs, _p), _ptr(dtypes), _ptr(tags), _ptr(cat_n), _ptr(labels), len(class_sorted),
            C.byref(h),
        )
    )
    dev = DeviceDataset.from_handle(h, device)
    try:
        width, stride = C.c_int32(0), C.c_int64(0)
        nuniq = np.zeros(nfeat, dtype=np.int32)
        off = np.zeros(nfeat + 1, dtype=np.int64)
        _check(load_library().rfmc_dataset_describe(h, C.byref(width), C.byref(stride), _ptr(nuniq), _ptr(off)))
        tables = np.zeros(max(1, int(off[-1])), dtype=np.float64)
        _check(load_library().rfmc_dataset_tables(h, _ptr(tables)))
        cols = []
        for j, name in enumerate(feature_cols):
            if j in cat_uniq:
                cols.append(ColumnCode(name, KIND_CATEGORICAL, cat_uniq[j], None, 0))
                continue
            table = tables[off[j] : off[j + 1]].copy()
            dt = dataset_numpy[name].dtype
            if dt.kind in "iu" and len(table):
                check_int_range(name, float(table[0]), float(table[-1]))
            cols.append(ColumnCode(name, KIND_NUMERIC, table.astype(dt), table, int(tags[j])))
    except Exception:
        dev.close()
        raise
    code_dtype = {1: np.uint8, 2: np.uint16, 4: np.uint32}[width.value]
    shape = (n, int(stride.value))

    def fetch_codes():
        out = np.empty(shape, dtype=code_dtype)
        _check(load_library().rfmc_dataset_codes(dev._h, _ptr(out)))
        return out

    enc = EncodedDataset(
        feature_cols=list(feature_cols),
        columns=cols,
        codes=None,
        width=int(width.value),
        stride=int(stride.value),
        labels=labels,
        class_sorted=class_sorted,
        class_vals=list(class_vals),
        sorted_to_classvals=s2c,
        class_rows=class_rows,
        fetch_codes=fetch_codes,
    )
    dev.enc = enc
    dev.set_class_rows(class_rows)  # the per-class row lists of the device draws, uploaded once with the dataset
    return enc, dev


class FlatFeatureLists:
    """The feature id lists of a plan's candidates stored back to back (what rfmc_batch_create takes): a
    read-only sequence of per-candidate arrays that costs the fit loop no Python work per candidate."""

    def __init__(self, flat: np.ndarray, counts: np.ndarray):
        self.flat = np.ascontiguousarray(flat)
        self.counts = np.ascontiguousarray(counts, dtype=np.int64).reshape(-1)
        self.off = np.zeros(len(self.counts) + 1, dtype=np.int64)
        np.cumsum(self.counts, out=self.off[1:])

    def __len__(self):
        return len(self.counts)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        return self.flat[self.off[i] : self.off[i + 1]]

    def __iter__(self):
        o = self.off.tolist()
        return (self.flat[a:b] for a, b in zip(o[:-1], o[1:]))


class GrowBatch:
    """rfmc_batch handle: grow + score a batch of candidates (plantTree + validationTree)."""

    def __init__(
        self,
        ds: DeviceDataset,
        idx_train,
        idx_val,
        cand_slot: Sequence[int],
        feat_lists: Sequence[Sequence[int]],
        max_depth: int,
        min_samples_split: int,
        stream=None,
        plan: Optional["DrawPlan"] = None,
    ):
        """Row ids either as host arrays ``idx_train`` / ``idx_val`` [n_slots, n], or -- ``plan`` given --
        the device-resident output of a DrawPlan on the same stream (both arrays None)."""
        self.ds = ds
        cand_slot = np.ascontiguousarray(cand_slot, dtype=np.int32)
        self.n_cand = len(cand_slot)
        if isinstance(feat_lists, FlatFeatureLists):
            off = feat_lists.off.astype(np.int32)
            flat = feat_lists.flat.astype(np.uint16)
        else:
            off = np.zeros(self.n_cand + 1, dtype=np.int32)
            off[1:] = np.cumsum([len(f) for f in feat_lists])
            flat = np.ascontiguousarray(np.concatenate([np.asarray(f, dtype=np.uint16) for f in feat_lists]), dtype=np.uint16)
        self._stream = _stream_ptr(stream)
        h = _p()
        if plan is not None:
            self._plan = plan  # the plan owns the row ids: keep it alive
            self.n_slots, self.n_train, self.n_val = plan.n_slots, plan.n_classes * plan.n_tr, plan.n_classes * plan.n_va
            _check(
                load_library().rfmc_batch_create_from_plan(
                    ds._h, plan._h, self.n_cand, _ptr(cand_slot), _ptr(off), _ptr(flat), int(min(max_depth, 2**31 - 1)),
                    int(min_samples_split), self._stream, C.byref(h),
                )
            )
        else:
            idx_train = np.ascontiguousarray(idx_train, dtype=np.int32)
            idx_val = np.ascontiguousarray(idx_val, dtype=np.int32)
            if idx_train.ndim != 2 or idx_val.ndim != 2 or idx_train.shape[0] != idx_val.shape[0]:
                raise ValueError("idx_train / idx_val must be [n_slots, n] arrays")
            self.n_slots, self.n_train = idx_train.shape
            self.n_val = idx_val.shape[1]
            _check(
                load_library().rfmc_batch_create(
                    ds._h, self.n_slots, self.n_train, self.n_val, _ptr(idx_train), _ptr(idx_val), self.n_cand,
                    _ptr(cand_slot), _ptr(off), _ptr(flat), int(min(max_depth, 2**31 - 1)), int(min_samples_split),
                    self._stream, C.byref(h),
                )
            )
        self._h = h
        self.correct = None
        self.n_nodes = None

    def run(self, stream=None):
        if stream is not None:
            self._stream = _stream_ptr(stream)
        _check(load_library().rfmc_batch_run(self._h, self._stream))
        return self

    def results(self):
        correct = np.zeros(self.n_cand, dtype=np.int32)
        n_nodes = np.zeros(self.n_cand, dtype=np.int32)
        _check(load_library().rfmc_batch_results(self._h, _ptr(correct), _ptr(n_nodes)))
        self.correct, self.n_nodes = correct, n_nodes
        return correct, n_nodes

    @property
    def launches(self) -> int:
        return load_library().rfmc_batch_launches(self._h)

    def set_timing(self, enabled: bool = True) -> None:
        """Measurement hook: bracket the kernels of every following run() with CUDA events."""
        _check(load_library().rfmc_batch_set_timing(self._h, int(bool(enabled))))

    def kernel_ms(self):
        """(slot gather ms, grower ms) of the last run(); waits for it."""
        g, k = C.c_float(0), C.c_float(0)
        _check(load_library().rfmc_batch_kernel_ms(self._h, C.byref(g), C.byref(k)))
        return float(g.value), float(k.value)

    def fetch(self, cand_ids: Sequence[int]) -> List[TreeSoA]:
        """Device -> host copy of the node tables of the selected candidates."""
        if self.n_nodes is None:
            self.results()
        ids = np.ascontiguousarray(cand_ids, dtype=np.int32)
        if len(ids) == 0:
            return []
        sizes = self.n_nodes[ids].astype(np.int64)
        total = int(sizes.sum())
        ncls = self.ds.enc.n_classes
        nodes = np.empty(total, dtype=NODE_DTYPE)
        counts = np.empty((total, ncls), dtype=np.int32)
        _check(load_library().rfmc_batch_pack(self._h, _ptr(ids), len(ids), _ptr(nodes), _ptr(counts), 0, self._stream))
        out, at = [], 0
        for s in sizes:
            s = int(s)
            out.append(TreeSoA(nodes[at : at + s], counts[at : at + s]))
            at += s
        return out

    def pack_to_device(self, cand_ids: Sequence[int], nodes_ptr: int, counts_ptr: int):
        """Pack selected candidates into caller-owned DEVICE buffers (e.g. an all-gather buffer)."""
        ids = np.ascontiguousarray(cand_ids, dtype=np.int32)
        _check(load_library().rfmc_batch_pack(self._h, _ptr(ids), len(ids), _p(nodes_ptr), _p(counts_ptr), 1, self._stream))

    def close(self):
        if getattr(self, "_h", None):
            load_library().rfmc_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GatheredSurvivors:
    """rfmc_gathered handle: every rank's surviving trees after the NCCL exchange (fitParallel across
    GPUs, model.py:403).  The node tables stay in HBM until ``tables(r)`` is called."""

    def __init__(self, comm_ptr: int, device: int, nodes_ptr: int, counts_ptr: int, sizes: np.ndarray, scores: np.ndarray,
                 masks: Optional[np.ndarray], n_classes: int, stream=None):
        sizes = np.ascontiguousarray(sizes, dtype=np.int64)
        scores = np.ascontiguousarray(scores, dtype=np.float64)
        masks = None if masks is None else np.ascontiguousarray(masks, dtype=np.uint8)
        self.n_classes = int(n_classes)
        self.mask_bytes = 0 if masks is None else int(masks.shape[1]) if masks.ndim == 2 else 0
        h = _p()
        _check(load_library().rfmc_allgather_survivors(
            _p(comm_ptr), int(device), _p(nodes_ptr) if nodes_ptr else None, _p(counts_ptr) if counts_ptr else None, int(sizes.sum()),
            len(sizes), _ptr(sizes), _ptr(scores), _ptr(masks) if self.mask_bytes else None, self.mask_bytes, self.n_classes,
            _stream_ptr(stream), C.byref(h)))
        self._h = h
        w, r, mt, mn = C.c_int32(0), C.c_int32(0), C.c_int64(0), C.c_int64(0)
        _check(load_library().rfmc_gathered_describe(h, C.byref(w), C.byref(r), C.byref(mt), C.byref(mn), None, None))
        self.world, self.rank = w.value, r.value
        self.trees_per_rank = np.zeros(self.world, dtype=np.int64)
        self.nodes_per_rank = np.zeros(self.world, dtype=np.int64)
        _check(load_library().rfmc_gathered_describe(h, None, None, None, None, _ptr(self.trees_per_rank), _ptr(self.nodes_per_rank)))

    def meta(self, r: int):
        """(tree sizes, survived scores, used-feature masks) of rank ``r``."""
        nt = int(self.trees_per_rank[r])
        sizes, scores = np.zeros(nt, dtype=np.int64), np.zeros(nt, dtype=np.float64)
        masks = np.zeros((nt, self.mask_bytes), dtype=np.uint8)
        _check(load_library().rfmc_gathered_meta(self._h, int(r), _ptr(sizes), _ptr(scores), _ptr(masks) if self.mask_bytes and nt else None))
        return sizes, scores, masks

    def tables(self, r: int):
        """Device -> host copy of rank ``r``'s node and class-count tables."""
        n = int(self.nodes_per_rank[r])
        nodes = np.empty(n, dtype=NODE_DTYPE)
        counts = np.empty((n, self.n_classes), dtype=np.int32)
        _check(load_library().rfmc_gathered_fetch(self._h, int(r), _ptr(nodes), _ptr(counts)))
        return nodes, counts

    def checksum(self) -> int:
        out = C.c_uint64(0)
        _check(load_library().rfmc_gathered_checksum(self._h, C.byref(out)))
        return int(out.value)

    def close(self):
        if getattr(self, "_h", None):
            load_library().rfmc_gathered_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceForest:
    """rfmc_forest handle: a flattened forest resident in HBM (useForest / testForest*)."""

    def __init__(self, tables, device: int = 0):
        require_gpu()
        self.tables = tables
        self.device = device
        self.n_classes = tables.n_classes
        h = _p()
        _check(
            load_library().rfmc_forest_create(
                device, _ptr(tables.pnodes), len(tables.pnodes), _ptr(tables.roots), len(tables.roots),
                len(tables.vote), _ptr(tables.vote), _ptr(tables.vote_absent), _ptr(tables.probs),
                _ptr(tables.probs_absent), tables.n_classes, _ptr(tables.scores), float(tables.score_fsum), C.byref(h),
            )
        )
        self._h = h
        nf = tables.n_features
        books = [t if t is not None else np.zeros(0) for t in tables.thresholds]
        off = np.zeros(nf + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(t) for t in books])
        flat = np.ascontiguousarray(np.concatenate(books) if books else np.zeros(0), dtype=np.float64)
        if len(flat) == 0:
            flat = np.zeros(1, dtype=np.float64)
        _check(load_library().rfmc_forest_set_codebook(self._h, nf, _ptr(off), _ptr(flat)))

    _DT = {"float64": 0, "float32": 1, "int64": 2, "int32": 3, "int16": 4, "int8": 5, "uint8": 6, "uint16": 7,
           "uint32": 8, "uint64": 9}

    @property
    def small_tree_stages(self) -> int:
        """> 0: the forest votes with the small-tree kernel (whole trees staged in shared memory)."""
        n = C.c_int32(0)
        _check(load_library().rfmc_forest_layout(self._h, C.byref(n)))
        return int(n.value)

    @property
    def vote_kernel(self) -> str:
        return "rfmc::predict_small_kernel" if self.small_tree_stages > 0 and self.tables.width <= 2 else "rfmc::predict_kernel"

    def predict_frame(self, frame, soft: bool, weighted: bool, want_probs: bool = True, want_labels: bool = True):
        """Vote a raw pandas frame: numeric columns go to the GPU as they are and are encoded there
        against the forest's split values; categorical columns are mapped to ids on the host.
        Mirrors flat.encode_frame + predict (tree.py:166-181 semantics for absent / NaN cells)."""
        import pandas as pd

        t = self.tables
        n = len(frame)
        absent = np.zeros(t.n_features, dtype=np.uint8)
        have = set(frame.columns)
        def categorical_ids(col, cats):
            # Series.factorize keeps Arrow-backed strings in Arrow (dictionary encode, GIL released)
            ids, uniques = col.factorize(use_na_sentinel=True)
            lut = np.zeros(len(uniques) + 1, dtype=np.int32)  # last slot: NaN sentinel (-1) -> 0
            for k, u in enumerate(uniques):
                try:
                    lut[k] = cats.get(u, 0)
                except TypeError:
                    lut[k] = 0
            return np.ascontiguousarray(lut[ids], dtype=np.int32)

        jobs = []  # (feature index, dtype code or None, array or column)
        for j, name in enumerate(t.feature_cols):
            thr, cats = t.thresholds[j], t.categories[j]
            if thr is None and cats is None:
                continue
            if name not in have:
                absent[j] = 1
                continue
            col = frame[name]
            if thr is not None:
                arr = col.to_numpy()
                if arr.dtype.name not in self._DT:
                    arr = pd.to_numeric(col, errors="coerce").to_numpy(dtype=np.float64, na_value=np.nan)
                arr = np.ascontiguousarray(arr)
                jobs.append((j, self._DT[arr.dtype.name], arr))
            else:
                jobs.append((j, None, (col, cats)))
        todo = [k for k, job in enumerate(jobs) if job[1] is None]
        if len(todo) > 1 and n >= 100_000:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(max_workers=min(8, len(todo))) as pool:
                done = list(pool.map(lambda k: categorical_ids(*jobs[k][2]), todo))
        else:
            done = [categorical_ids(*jobs[k][2]) for k in todo]
        for k, arr in zip(todo, done):
            jobs[k] = (jobs[k][0], 10, arr)
        keep = [job[2] for job in jobs]
        ptrs = [a.ctypes.data for a in keep]
        feats = [job[0] for job in jobs]
        dts = [job[1] for job in jobs]
        probs = np.empty((n, self.n_classes), dtype=np.float64) if want_probs else None
        labels = np.empty(n, dtype=np.int32) if want_labels else None
        ncol = len(keep)
        ptr_arr = (C.c_void_p * max(ncol, 1))(*ptrs)
        feat_arr = np.asarray(feats if feats else [0], dtype=np.int32)
        dt_arr = np.asarray(dts if dts else [0], dtype=np.int32)
        _check(
            load_library().rfmc_forest_predict_frame(
                self._h, n, t.n_features, t.stride, t.width, ncol, C.cast(ptr_arr, _p), _ptr(feat_arr), _ptr(dt_arr),
                _ptr(absent), int(bool(soft)), int(bool(weighted)), _ptr(probs), _ptr(labels),
            )
        )
        return probs, labels

    def predict(self, codes: np.ndarray, absent: Optional[np.ndarray], soft: bool, weighted: bool,
                want_probs: bool = True, want_labels: bool = True, stream=None):
        codes = np.ascontiguousarray(codes)
        n, stride = codes.shape
        probs = np.empty((n, self.n_classes), dtype=np.float64) if want_probs else None
        labels = np.empty(n, dtype=np.int32) if want_labels else None
        ab = None if absent is None else np.ascontiguousarray(absent, dtype=np.uint8)
        _check(
            load_library().rfmc_forest_predict(
                self._h, _ptr(codes), codes.dtype.itemsize, n, stride, self.tables.n_features, _ptr(ab), int(bool(soft)),
                int(bool(weighted)), _ptr(probs), _ptr(labels), 0, _stream_ptr(stream),
            )
        )
        return probs, labels

    def leaves(self, codes: np.ndarray, absent: Optional[np.ndarray], stream=None):
        """Per-(row, tree) payload index and crossed-an-absent-column flag."""
        codes = np.ascontiguousarray(codes)
        n, stride = codes.shape
        out = np.empty((n, self.tables.n_trees), dtype=np.int32)
        ab = None if absent is None else np.ascontiguousarray(absent, dtype=np.uint8)
        _check(
            load_library().rfmc_forest_leaves(
                self._h, _ptr(codes), codes.dtype.itemsize, n, stride, self.tables.n_features, _ptr(ab), _ptr(out),
                _stream_ptr(stream),
            )
        )
        payload = out & 0x7FFFFFFF
        return payload, out < 0

    def predict_device(self, codes_ptr: int, width: int, n_rows: int, stride: int, absent: Optional[np.ndarray],
                       soft: bool, weighted: bool, probs_ptr: int, labels_ptr: int, stream=None):
        """Device-resident variant: pointers are CUDA device addresses; asynchronous on ``stream``."""
        ab = None if absent is None else np.ascontiguousarray(absent, dtype=np.uint8)
        _check(
            load_library().rfmc_forest_predict(
                self._h, _p(codes_ptr), width, n_rows, stride, self.tables.n_features, _ptr(ab), int(bool(soft)),
                int(bool(weighted)), _p(probs_ptr) if probs_ptr else None, _p(labels_ptr) if labels_ptr else None, 1,
                _stream_ptr(stream),
            )
        )

    def close(self):
        if getattr(self, "_h", None):
            load_library().rfmc_forest_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
