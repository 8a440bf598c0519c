"""GPU (B200): the grower/scorer against the CPU code model over the kernel's dispatch space --
code widths (uint8 / uint16 / uint32), position widths (uint16 / uint32 rows per candidate),
shared-memory vs global-memory position arrays, class counts up to 40, single-feature lists,
depth and min_samples_split limits, empty hold-out sets, tiny bootstrap sets."""
import numpy as np
import pandas as pd
import pytest

from oracle import code_model as cm
from oracle import encode_ref as ref_enc

pytestmark = pytest.mark.gpu


def _frame(seed, n, n_classes, kind):
    g = np.random.default_rng(seed)
    if kind == "u8":  # every column <= 255 distinct values -> uint8 codes
        cols = {
            "a": g.integers(0, 200, size=n).astype(np.int16),
            "b": np.round(g.normal(size=n), 1),
            "c": g.choice(list("pqrstu"), size=n),
            "d": g.integers(0, 2, size=n).astype(np.uint8),
        }
    elif kind == "u32":  # a raw float column: > 65535 distinct values -> uint32 codes
        cols = {
            "a": g.normal(size=n),
            "b": g.integers(0, 100000, size=n),
            "c": g.choice([f"k{k}" for k in range(40)], size=n),
            "d": g.normal(size=n).astype(np.float32),
        }
    else:
        cols = {
            "a": np.round(g.normal(size=n), 3),
            "b": g.integers(0, 3000, size=n),
            "c": g.choice([f"k{k}" for k in range(15)], size=n),
            "d": np.round(g.normal(size=n), 2).astype(np.float32),
            "e": g.integers(-5, 5, size=n).astype(np.int8),
        }
    df = pd.DataFrame(cols)
    first = list(cols)[0]
    s = df[first].astype(float) / (df[first].astype(float).std() + 1e-9) + g.normal(size=n) * 0.8
    df["target"] = pd.qcut(s, n_classes, labels=[f"c{k:02d}" for k in g.permutation(n_classes)]).astype(str)
    return df


CASES = [
    # seed, rows, classes, kind, n_train, n_val, mss, max_depth, n_cand
    (1, 4000, 3, "u8", 300, 50, 1, 1000, 6),
    (2, 120000, 4, "u32", 1200, 200, 1, 1000, 4),
    (3, 6000, 40, "u16", 2000, 300, 1, 1000, 4),
    (4, 6000, 2, "u16", 1, 5, 1, 1000, 3),          # one-row bootstrap: root leaf
    (5, 6000, 3, "u16", 4, 0, 1, 1000, 5),          # tiny root, empty hold-out set
    (6, 6000, 3, "u16", 37, 9, 2, 4, 5),
    (7, 6000, 5, "u16", 900, 100, 7, 1000, 4),      # min_samples_split
    (8, 6000, 5, "u16", 900, 100, 1, 1, 3),         # max_depth 1: root leaf with mixed classes
    (9, 90000, 4, "u16", 24000, 500, 1, 1000, 2),   # positions no longer fit shared memory
    (10, 90000, 3, "u16", 70000, 1000, 1, 1000, 1), # > 65535 rows per candidate: uint32 positions
]


@pytest.mark.parametrize("seed,n,n_classes,kind,n_train,n_val,mss,max_depth,n_cand", CASES)
def test_dispatch_space_matches_code_model(seed, n, n_classes, kind, n_train, n_val, mss, max_depth, n_cand):
    from random_forest_mc_b200 import engine as eng

    eng.require_gpu()
    df = _frame(seed, n, n_classes, kind)
    feats_all = [c for c in df.columns if c != "target"]
    numeric = df.select_dtypes([np.number]).columns.to_list()
    cols = {c: df[c].to_numpy() for c in df.columns}
    enc = ref_enc.encode_dataset(cols, feats_all, numeric, "target", df["target"].unique().tolist())
    assert enc.width == {"u8": 1, "u16": 2, "u32": 4}[kind]
    g = np.random.default_rng(seed + 500)
    idx_T = np.stack([g.choice(n, n_train, replace=False) for _ in range(2)])
    idx_V = np.stack([g.choice(n, n_val, replace=False) for _ in range(2)]) if n_val else np.zeros((2, 0), dtype=np.int64)
    slot_of, feats = [], []
    for i in range(n_cand):
        k = 1 if i == 0 else int(g.integers(1, len(feats_all) + 1))
        feats.append([int(x) for x in g.permutation(len(feats_all))[:k]])
        slot_of.append(i % 2)
    ds = eng.DeviceDataset(enc)
    b = eng.GrowBatch(ds, idx_T, idx_V, slot_of, feats, max_depth, mss).run()
    correct, n_nodes = b.results()
    trees = b.fetch(list(range(n_cand)))
    for i in range(n_cand):
        nodes, counts = cm.grow_codes(enc, idx_T[slot_of[i]], feats[i], max_depth, mss)
        assert trees[i].n_nodes == len(nodes)
        for k in ("feat", "thr", "child", "flags", "sval"):
            assert np.array_equal(trees[i].nodes[k], nodes[k]), (i, k)
        assert np.array_equal(trees[i].counts, counts)
        if n_val:
            assert correct[i] == cm.score_codes(enc, nodes, counts, idx_V[slot_of[i]])
    b.close()
    ds.close()
