"""GPU parity tests proper (B200): the CUDA kernels, called through the C ABI, against the CPU
oracle and the vectors recorded from the reference.  Bit-exact everywhere (integer/index work and
fp64 sums in the reference's order)."""
import numpy as np
import pandas as pd
import pytest

from oracle import code_model as cm
from oracle import encode_ref as ref_enc
from random_forest_mc_b200.flat import TreeSoA, encode_frame, forest_tables_from_dicts, soa_to_dict
from tests.helpers import golden_names, load_golden, tree_diff
from tests.test_code_model import encode_golden

pytestmark = pytest.mark.gpu


def _engine():
    from random_forest_mc_b200 import engine

    engine.require_gpu()
    return engine


def _same_tables(a: TreeSoA, nodes, counts):
    assert a.n_nodes == len(nodes)
    for k in ("feat", "thr", "child", "flags"):
        assert np.array_equal(a.nodes[k], nodes[k]), k
    assert np.array_equal(a.nodes["sval"], nodes["sval"])
    assert np.array_equal(a.counts, counts)


@pytest.mark.parametrize("name", golden_names())
def test_grow_and_score_match_reference_vectors(name):
    eng = _engine()
    g = load_golden(name)
    enc = encode_golden(g)
    kw = g["kwargs"]
    max_depth = kw.get("max_depth") or 1000
    mss = kw.get("min_samples_split", 1)
    fid = {n: i for i, n in enumerate(g["feature_cols"])}
    ds = eng.DeviceDataset(enc)
    # one slot per distinct (idx_T, idx_V) pair, one candidate per recorded trace entry
    slots, slot_of = [], []
    for ref in g["trace"]:
        key = (ref["idx_T"].tobytes(), ref["idx_V"].tobytes())
        if not slots or slots[-1][0] != key:
            slots.append((key, ref["idx_T"], ref["idx_V"]))
        slot_of.append(len(slots) - 1)
    idx_T = np.stack([s[1] for s in slots])
    idx_V = np.stack([s[2] for s in slots])
    feats = [[fid[f] for f in ref["F"]] for ref in g["trace"]]
    b = eng.GrowBatch(ds, idx_T, idx_V, slot_of, feats, max_depth, mss)
    b.run()
    correct, n_nodes = b.results()
    trees = b.fetch(list(range(len(feats))))
    for i, ref in enumerate(g["trace"]):
        nodes, counts = cm.grow_codes(enc, ref["idx_T"], feats[i], max_depth, mss)
        _same_tables(trees[i], nodes, counts)
        assert tree_diff(soa_to_dict(trees[i], enc, g["type_of_cols"]), ref["tree"]) is None
        assert np.float64(correct[i]) / len(ref["idx_V"]) == float(ref["th_val"])
    b.close()
    ds.close()


def _synthetic(seed, n, high_card=False):
    g = np.random.default_rng(seed)
    cols = {
        "a": np.round(g.normal(size=n), 3),
        "b": g.integers(0, 1000, size=n),
        "c": np.round(g.normal(size=n), 2).astype(np.float32),
        "d": g.choice([f"v{k}" for k in range(9)], size=n),
        "e": g.integers(0, 3, size=n).astype(np.int8),
        "f": g.normal(size=n),
    }
    if high_card:
        cols["g"] = g.choice([f"id{k}" for k in range(700)], size=n)
        cols["h"] = g.integers(-(2**31), 2**31 - 1, size=n).astype(np.int32)
    df = pd.DataFrame(cols)
    s = df["a"] + df["b"] / 500 + (df["d"] == "v1") + g.normal(size=n) * 0.7
    df["target"] = pd.qcut(s, 4, labels=["3", "0", "2", "1"]).astype(str)
    return df


@pytest.mark.parametrize("seed,n,n_train,n_val,high_card,mss,max_depth", [
    (11, 20000, 400, 100, False, 1, 1000),
    (12, 20000, 3000, 500, False, 1, 1000),
    (13, 30000, 2500, 300, True, 1, 1000),
    (14, 20000, 1500, 200, True, 4, 9),
    (15, 5000, 33, 7, False, 1, 1000),
])
def test_grow_matches_code_model_on_synthetic(seed, n, n_train, n_val, high_card, mss, max_depth):
    eng = _engine()
    df = _synthetic(seed, n, high_card)
    feats_all = [c for c in df.columns if c != "target"]
    numeric = df.select_dtypes([np.number]).columns.to_list()
    cols = {c: df[c].to_numpy() for c in df.columns}
    enc = ref_enc.encode_dataset(cols, feats_all, numeric, "target", df["target"].unique().tolist())
    g = np.random.default_rng(seed + 1000)
    n_slots, per_slot = 3, 3
    idx_T = np.stack([g.choice(n, n_train, replace=False) for _ in range(n_slots)])
    idx_V = np.stack([g.choice(n, n_val, replace=False) for _ in range(n_slots)])
    slot_of, feats = [], []
    for s in range(n_slots):
        for _ in range(per_slot):
            k = int(g.integers(2, len(feats_all) + 1))
            feats.append([int(x) for x in g.permutation(len(feats_all))[:k]])
            slot_of.append(s)
    ds = eng.DeviceDataset(enc)
    b = eng.GrowBatch(ds, idx_T, idx_V, slot_of, feats, max_depth, mss)
    b.run()
    correct, n_nodes = b.results()
    trees = b.fetch(list(range(len(feats))))
    for i in range(len(feats)):
        nodes, counts = cm.grow_codes(enc, idx_T[slot_of[i]], feats[i], max_depth, mss)
        _same_tables(trees[i], nodes, counts)
        assert correct[i] == cm.score_codes(enc, nodes, counts, idx_V[slot_of[i]])
    b.close()
    ds.close()


@pytest.mark.parametrize("name", golden_names())
def test_voting_matches_reference_vectors(name):
    eng = _engine()
    g = load_golden(name)
    t = forest_tables_from_dicts(g["forest"], g["survived_scores"], g["class_vals"], g["feature_cols"])
    forest = eng.DeviceForest(t)
    for (fname, soft, weighted), want in g["preds"].items():
        codes, absent = encode_frame(t, g["frames"][fname])
        probs, labels = forest.predict(codes, absent, soft, weighted)
        ref = np.array([[d[c] for c in g["class_vals"]] for d in want["probs"]], dtype=np.float64)
        assert np.array_equal(probs, ref), (fname, soft, weighted)
        assert [g["class_vals"][k] for k in labels] == want["labels"], (fname, soft, weighted)
    forest.close()


def test_voting_many_rows_matches_cpu_model():
    """More rows than one CTA, ragged last tile, every mode; checked against the CPU statement."""
    eng = _engine()
    g = load_golden("mixed_big")
    t = forest_tables_from_dicts(g["forest"], g["survived_scores"], g["class_vals"], g["feature_cols"])
    frame = g["frame"].drop(columns=["target"]).head(777)
    codes, absent = encode_frame(t, frame)
    forest = eng.DeviceForest(t)
    for soft in (False, True):
        for weighted in (False, True):
            probs, labels = forest.predict(codes, absent, soft, weighted)
            p2, l2 = cm.vote_tables(t, codes, absent, soft, weighted)
            assert np.array_equal(probs, p2)
            assert np.array_equal(labels, l2)
    forest.close()


@pytest.mark.parametrize("name", golden_names())
def test_small_tree_and_generic_vote_kernels_agree(name, monkeypatch):
    """Golden forests are small: they vote with predict_small_kernel (whole trees in shared memory, fixed-depth
    branch-free walks).  The generic blocked walk (RFMC_PRED_SMALL=0) must give the same bits in every mode,
    with missing cells and a ragged last tile; both equal the CPU statement."""
    eng = _engine()
    g = load_golden(name)
    t = forest_tables_from_dicts(g["forest"], g["survived_scores"], g["class_vals"], g["feature_cols"])
    forest = eng.DeviceForest(t)
    assert forest.small_tree_stages > 0 and forest.vote_kernel.endswith("predict_small_kernel")
    frame = g["frames"]["nan"] if "nan" in g["frames"] else g["frames"]["clean"]
    reps = max(1, 600 // max(1, len(frame)))
    codes, absent = encode_frame(t, frame)
    codes = np.ascontiguousarray(np.tile(codes, (reps, 1)))[: max(len(frame), 517)]
    assert absent is None or not np.any(absent)
    for soft in (False, True):
        for weighted in (False, True):
            monkeypatch.delenv("RFMC_PRED_SMALL", raising=False)
            monkeypatch.delenv("RFMC_PRED_SMALL_REC4", raising=False)
            p_small, l_small = forest.predict(codes, absent, soft, weighted)  # 4-byte records when the forest allows them
            monkeypatch.setenv("RFMC_PRED_SMALL_REC4", "0")
            p_rec8, l_rec8 = forest.predict(codes, absent, soft, weighted)  # 8-byte records
            monkeypatch.setenv("RFMC_PRED_SMALL", "0")
            p_gen, l_gen = forest.predict(codes, absent, soft, weighted)
            p_cpu, l_cpu = cm.vote_tables(t, codes, absent, soft, weighted)
            assert np.array_equal(p_small, p_gen) and np.array_equal(l_small, l_gen), (soft, weighted)
            assert np.array_equal(p_rec8, p_gen) and np.array_equal(l_rec8, l_gen), (soft, weighted)
            assert np.array_equal(p_small, p_cpu) and np.array_equal(l_small, l_cpu), (soft, weighted)
    monkeypatch.delenv("RFMC_PRED_SMALL", raising=False)
    monkeypatch.delenv("RFMC_PRED_SMALL_REC4", raising=False)
    forest.close()


def test_cache_trim_gives_the_pools_back_and_the_engine_keeps_working():
    """rfmc_cache_trim frees the cached device blocks / pinned buffers of closed batches; the next batch
    allocates afresh and gives the same trees."""
    eng = _engine()
    g = load_golden("mixed_default")
    enc = encode_golden(g)
    ds = eng.DeviceDataset(enc)
    ref = g["trace"][0]
    fidx = {name: j for j, name in enumerate(g["feature_cols"])}
    feats = [[fidx[f] for f in ref["F"]]]

    def grow():
        b = eng.GrowBatch(ds, np.asarray(ref["idx_T"])[None, :], np.asarray(ref["idx_V"])[None, :], [0], feats, 1000, g["kwargs"].get("min_samples_split", 1))
        b.run()
        correct, _ = b.results()
        tree = b.fetch([0])[0]
        b.close()
        return int(correct[0]), tree

    c1, t1 = grow()
    freed = eng.cache_trim(0)
    assert freed > 0
    assert eng.cache_trim(0) == 0
    c2, t2 = grow()
    assert c1 == c2 and np.array_equal(t1.nodes, t2.nodes) and np.array_equal(t1.counts, t2.counts)
    ds.close()
