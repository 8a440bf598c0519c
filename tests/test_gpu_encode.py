"""GPU (B200): `rfmc_dataset_encode` -- process_dataset's rank coding on the device -- against the
numpy restatement (oracle/encode_ref.py): code matrix bit-equal, same code width and row pitch, same
value tables (dtype included), on the reference's recorded frames and on columns that exercise every
dtype, signed zeros, infinities, full-range integers, the uint32 code path, a 700-level categorical
and sizes on both sides of the shared-memory sort tile (global bitonic passes)."""
import numpy as np
import pandas as pd
import pytest

from oracle import encode_ref
from tests.helpers import golden_names, load_golden

pytestmark = pytest.mark.gpu


def both(cols, feats, numeric, target="target"):
    from random_forest_mc_b200 import engine as eng

    class_vals = pd.Series(cols[target]).unique().tolist()
    ref = encode_ref.encode_dataset(cols, feats, numeric, target, class_vals)
    enc, dev = eng.encode_dataset_device(cols, feats, numeric, target, class_vals)
    return ref, enc, dev


def assert_same(ref, enc):
    assert enc.width == ref.width and enc.stride == ref.stride
    assert enc.codes.dtype == ref.codes.dtype and np.array_equal(enc.codes, ref.codes)
    assert np.array_equal(enc.labels, ref.labels) and enc.class_sorted == ref.class_sorted
    for a, b in zip(enc.columns, ref.columns):
        assert a.kind == b.kind and a.tag == b.tag and a.n_unique == b.n_unique, a.name
        assert a.uniq.dtype == b.uniq.dtype and np.array_equal(a.uniq, b.uniq), a.name
        if b.table is not None:
            assert np.array_equal(a.table, b.table), a.name
            assert not np.signbit(a.table[a.table == 0]).any()  # zero is always +0.0


@pytest.mark.parametrize("name", golden_names())
def test_recorded_frames(name):
    g = load_golden(name)
    frame = g["frame"].copy()
    tcol = g["kwargs"]["target_col"]
    frame[tcol] = frame[tcol].astype(str)
    frame = frame.dropna()
    cols = {c: frame[c].to_numpy() for c in frame.columns}
    ref, enc, dev = both(cols, g["feature_cols"], g["numeric_cols"], tcol)
    assert_same(ref, enc)
    dev.close()


@pytest.mark.parametrize("n", [1, 2, 37, 4096, 4097, 8192, 20_001, 300_000])
def test_every_dtype(n):
    r = np.random.default_rng(n)
    cols, numeric = {}, []
    for dt in ("f8", "f4", "i8", "i4", "i2", "i1", "u1", "u2", "u4", "u8"):
        info = np.iinfo(dt) if dt[0] in "iu" else None
        if info is not None:
            lo, hi = max(info.min, -(2**51)), min(info.max, 2**51)
            v = r.integers(lo, hi, size=n, endpoint=True).astype(dt)
            w = r.integers(max(lo, -50), min(hi, 50), size=n, endpoint=True).astype(dt)
        else:
            v = (r.normal(size=n) * 10.0 ** r.integers(-3, 6, size=n)).astype(dt)
            w = np.round(r.normal(size=n), 1).astype(dt)
            w[r.random(n) < 0.1] = -0.0
            w[r.random(n) < 0.1] = 0.0
            if n > 8:
                w[:2] = [np.inf, -np.inf]
        cols[f"wide_{dt}"], cols[f"few_{dt}"] = v, w
        numeric += [f"wide_{dt}", f"few_{dt}"]
    levels = np.array([f"v{k}" for k in range(700)], dtype=object)
    cols["cat"] = levels[r.integers(0, min(700, n), size=n)]
    cols["flag"] = r.choice(np.array([True, False], dtype=object), size=n)
    cols["target"] = np.array([str(k) for k in r.integers(0, 3, size=n)], dtype=object)
    feats = [c for c in cols if c != "target"]
    ref, enc, dev = both(cols, feats, numeric)
    assert_same(ref, enc)
    dev.close()


def test_narrow_codes_and_strided_input():
    r = np.random.default_rng(5)
    n = 5000
    block = r.integers(0, 200, size=(n, 3)).astype(np.int64)  # columns of a 2-D block are strided views
    cols = {"a": block[:, 0], "b": block[:, 1], "c": block[:, 2].astype(">i4"), "target": np.array(["x", "y"], dtype=object)[r.integers(0, 2, n)]}
    ref, enc, dev = both(cols, ["a", "b", "c"], ["a", "b", "c"])
    assert enc.width == 1
    assert_same(ref, enc)
    dev.close()


def test_unsupported_inputs_raise():
    from random_forest_mc_b200 import engine as eng

    n = 100
    y = np.array(["0", "1"], dtype=object)[np.arange(n) % 2]
    with pytest.raises(NotImplementedError):
        eng.encode_dataset_device({"a": np.arange(n, dtype=np.float16), "target": y}, ["a"], ["a"], "target", ["0", "1"])
    big = np.arange(n, dtype=np.int64)
    big[3] = 2**53
    with pytest.raises(NotImplementedError):
        eng.encode_dataset_device({"a": big, "target": y}, ["a"], ["a"], "target", ["0", "1"])


def test_process_dataset_full_size_matches_restatement():
    """BASELINE C2 shape (1M x 64): the model's process_dataset (device coding) == numpy restatement."""
    import logging

    import bench
    from random_forest_mc_b200.model import RandomForestMC

    logging.disable(logging.WARNING)
    frame = bench.make_frame(1_000_000, 64, 4, seed=0)
    m = RandomForestMC(n_trees=2, target_col="target", max_discard_trees=2)
    m.process_dataset(frame.copy())
    ref = encode_ref.encode_dataset(m.dataset_numpy, m.feature_cols, m.numeric_cols, "target", m.class_vals)
    assert_same(ref, m.encoded)
    for a, b in zip(m.encoded.class_rows, ref.class_rows):
        assert np.array_equal(a, b)
