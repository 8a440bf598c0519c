"""CPU: encode + code-domain restatement + SoA->dict conversion reproduce the reference's recorded
candidates (tree dicts incl. types/key order, hold-out scores)."""
import numpy as np
import pytest

from oracle import code_model as cm
from oracle import encode_ref as ref_enc
from random_forest_mc_b200 import encode as enc_mod
from random_forest_mc_b200.flat import TreeSoA, soa_to_dict
from tests.helpers import golden_names, load_golden, tree_diff


def encode_golden(g):
    frame = g["frame"].copy()
    tcol = g["kwargs"]["target_col"]
    frame[tcol] = frame[tcol].astype(str)
    frame = frame.dropna()
    cols = {c: frame[c].to_numpy() for c in frame.columns}
    return ref_enc.encode_dataset(cols, g["feature_cols"], g["numeric_cols"], tcol, g["class_vals"])


@pytest.mark.parametrize("name", golden_names())
def test_code_domain_growth_equals_reference(name):
    g = load_golden(name)
    enc = encode_golden(g)
    kw = g["kwargs"]
    max_depth = kw.get("max_depth") or 1000
    mss = kw.get("min_samples_split", 1)
    fid = {n: i for i, n in enumerate(g["feature_cols"])}
    for ref in g["trace"]:
        nodes, counts = cm.grow_codes(enc, ref["idx_T"], [fid[f] for f in ref["F"]], max_depth, mss)
        mine = soa_to_dict(TreeSoA(nodes, counts), enc, g["type_of_cols"])
        assert tree_diff(mine, ref["tree"]) is None
        correct = cm.score_codes(enc, nodes, counts, ref["idx_V"])
        assert np.float64(correct) / len(ref["idx_V"]) == float(ref["th_val"])


def test_codes_preserve_order_and_equality():
    g = load_golden("mixed_default")
    enc = encode_golden(g)
    frame = g["frame"]
    for j, col in enumerate(enc.columns):
        raw = frame[col.name].to_numpy()
        back = col.uniq[enc.codes[:, j].astype(np.int64) - 1]
        assert (back == raw).all()
        assert (enc.codes[:, j] >= 1).all()
    assert (enc.stride * enc.width) % 4 == 0 and ((enc.stride * enc.width) // 4) % 2 == 1


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int8, np.uint8, np.int16, np.int32, np.int64, np.uint32])
def test_median_value_equals_numpy_quantile(dtype):
    rng = np.random.default_rng(7)
    info = np.iinfo(dtype) if np.dtype(dtype).kind in "iu" else None
    for trial in range(400):
        n = int(rng.integers(3, 12))
        if info is not None:
            lo, hi = max(info.min, -(2**40)), min(info.max, 2**40)
            v = rng.integers(lo, hi, size=n, endpoint=True).astype(dtype)
            if trial % 3 == 0:
                v = rng.choice(np.array([lo, hi, 0, 1], dtype=dtype), size=n)
        else:
            v = (rng.normal(size=n) * 10.0 ** rng.integers(-3, 6)).astype(dtype)
            if trial % 5 == 0:
                v = np.round(v, 1)
        want = float(np.quantile(v, 0.5))
        s = np.sort(v)
        k0 = (n - 1) // 2
        cc, _ = ref_enc.encode_column("x", v, True)
        got = enc_mod.median_value(float(s[k0]), float(s[k0 + 1]), n % 2 == 1, cc.tag)
        assert got == want or (got != got and want != want), (v, got, want)
