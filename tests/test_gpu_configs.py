"""GPU (B200): BASELINE.json's configurations 3, 4 and 5 as parity cases.

C3  1M x 64 with missing values, weighted soft voting -- at full size.  Train frame: 20 % of the
    ROWS hold a NaN and are dropped (model.py:181); predict frames: 20 % of the CELLS are NaN (-> '<',
    tree.py:176-181) and, separately, 13 of 64 columns are absent (-> '>=' + renormalised first leaf,
    tree.py:166-175,192-210).  Probabilities are compared bit for bit with the oracle on a slice the
    oracle finishes in seconds, and through size-independent properties on the whole frame.
C4  sharded fit on 10M x 128 (96 numeric + 32 categorical), 2000 + 500 rows per class -- at full
    size when the box has the host memory for the pandas frame (>= 128 GB free), else 2M rows.
    SCALED: 2 x 4 slots instead of 8 x 256; the contract is SURVEY 8e: shard r == the serial
    algorithm under seed 1000 + r, forest = concatenation in shard order.  The NCCL exchange itself
    is covered by tests/test_gpu_dist.py.
C5  predict_proba over a forest merged from 16 models x 256 trees (4096 trees, mergeForest by="add").
    SCALED: 1M rows instead of 100M, voted as 8 contiguous row shards (the 8-GPU partition) and as
    one piece; oracle parity on a slice; merged hard votes == the sum of the 16 models' votes."""
import logging
import random

import numpy as np
import pandas as pd
import pytest

from tests.helpers import probs_equal

pytestmark = pytest.mark.gpu


def _probs_matrix(dicts, class_vals):
    return np.array([[d[c] for c in class_vals] for d in dicts], dtype=np.float64)


def test_c3_missing_values_weighted_soft_voting():
    import bench
    from oracle import rfmc_oracle as orc
    from random_forest_mc_b200.model import RandomForestMC

    logging.disable(logging.WARNING)
    r = np.random.default_rng(3)
    frame = bench.make_frame(1_000_000, 64, 4, seed=0)
    feats = [c for c in frame.columns if c != "target"]
    numeric = list(frame[feats].select_dtypes([np.number]).columns)
    floats = [c for c in numeric if frame[c].dtype.kind == "f"]
    holed = r.choice(len(frame), size=200_000, replace=False)  # 20 % of the rows get one NaN
    train = frame.copy()
    for k, c in enumerate(floats):
        train.loc[holed[k :: len(floats)], c] = np.nan
    m = RandomForestMC(n_trees=32, target_col="target", max_discard_trees=4)
    np.random.seed(3)
    random.seed(3)
    m.fit(train)
    assert m.n_samples == 800_000 and m.encoded.n_rows == 800_000 and len(m.data) == 32
    m.setSoftVoting(True)
    m.setWeightedTrees(True)
    trees, scores = [t.data for t in m.data], list(m.survived_scores)

    # 20 % of the cells missing
    nan_frame = frame.drop(columns=["target"]).iloc[:200_000].copy()
    for c in feats:
        hit = r.random(len(nan_frame)) < 0.2
        if c in numeric:
            col = nan_frame[c].astype("float64") if nan_frame[c].dtype.kind in "iu" else nan_frame[c].copy()
            col[hit] = np.nan
            nan_frame[c] = col
        else:
            nan_frame.loc[hit, c] = None
    got = m.predict_proba(nan_frame)
    want = orc.forest_probs(trees, scores, m.class_vals, nan_frame.iloc[:150], soft=True, weighted=True)
    assert probs_equal(got[:150], want)
    P = _probs_matrix(got, m.class_vals)
    assert np.all(np.abs(P.sum(axis=1) - 1.0) < 1e-12) and P.min() >= 0.0
    halves = m.predict_proba(nan_frame.iloc[:77_777]) + m.predict_proba(nan_frame.iloc[77_777:])
    assert probs_equal(got, halves)  # chunking / row sharding does not change a bit
    labels = m.predict(nan_frame)
    assert labels == [m.class_vals[k] for k in P.argmax(axis=1)]  # first maximum in class_vals order

    # 13 of 64 columns absent from the frame
    dropped = list(r.choice(feats, size=13, replace=False))
    absent_frame = frame.drop(columns=["target"] + dropped).iloc[:100_000]
    got = m.predict_proba(absent_frame)
    want = orc.forest_probs(trees, scores, m.class_vals, absent_frame.iloc[:150], soft=True, weighted=True)
    assert probs_equal(got[:150], want)
    P = _probs_matrix(got, m.class_vals)
    assert np.all(np.abs(P.sum(axis=1) - 1.0) < 1e-12)
    # hard voting on the same frames agrees with the oracle too
    m.setSoftVoting(False)
    got = m.predict_proba(absent_frame.iloc[:150])
    assert probs_equal(got, orc.forest_probs(trees, scores, m.class_vals, absent_frame.iloc[:150], soft=False, weighted=True))


def test_c4_shape_sharded_fit_contract():
    import bench
    from random_forest_mc_b200.model import RandomForestMC

    import psutil

    logging.disable(logging.WARNING)
    n_rows = 10_000_000 if psutil.virtual_memory().available >= (128 << 30) else 2_000_000
    frame = bench.make_frame(n_rows, 128, 4, seed=1)
    kw = dict(target_col="target", max_discard_trees=4, batch_train_pclass=2000, batch_val_pclass=500)
    shards = []
    m = RandomForestMC(n_trees=4, **kw)
    m.process_dataset(frame)
    assert m.encoded.n_features == 128 and m.encoded.n_rows == n_rows
    for rank in range(2):  # what rank r of a 2-GPU run computes (dist.fit_sharded), here one after the other
        m.reset_forest()
        np.random.seed(1000 + rank)
        random.seed(1000 + rank)
        m.fit(disable_progress_bar=True)
        shards.append([(t.soa.nodes.copy(), t.soa.counts.copy(), float(t.survived_score)) for t in m.data])
        assert len(m.data) == 4
    # the same shard through another launch cap is bit-identical (speculation + rewind are exact)
    m.reset_forest()
    m.max_candidates_per_launch = 3
    np.random.seed(1001)
    random.seed(1001)
    m.fit(disable_progress_bar=True)
    for (n1, c1, s1), t in zip(shards[1], m.data):
        assert s1 == float(t.survived_score) and np.array_equal(n1, t.soa.nodes) and np.array_equal(c1, t.soa.counts)
    # shards differ (different seeds) and every tree is a valid partition of its 8000 bootstrap rows
    assert not np.array_equal(shards[0][0][0], shards[1][0][0])
    for nodes, counts, score in shards[0] + shards[1]:
        leaf = nodes["feat"] < 0
        assert counts[0].sum() == 8000 and counts[leaf].sum() == 8000 and 0.0 <= score <= 1.0
    # a 128-feature candidate's hold-out score == the vote kernel on the same rows (two kernels agree)
    np.random.seed(7)
    random.seed(7)
    plan = m._draw_plan(1, 2)
    batch, _ = m._grow_plan(plan)
    correct, _ = batch.results()
    for k in range(2):
        tree = m._wrap(batch.fetch([k])[0])
        assert float(m.validationTree(tree, m._slot_rows(plan[0])[1])) == correct[k] / batch.n_val
    batch.close()


def test_c5_merged_forest_row_sharded_predict():
    import bench
    from oracle import rfmc_oracle as orc
    from random_forest_mc_b200.model import RandomForestMC

    logging.disable(logging.WARNING)
    frame = bench.make_frame(1_000_000, 64, 4, seed=0)
    models = []
    base = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4)
    base.process_dataset(frame.copy())
    for seed in range(16):
        np.random.seed(seed)
        random.seed(seed)
        base.reset_forest()
        base.fit(disable_progress_bar=True)
        part = RandomForestMC(n_trees=256, target_col="target", max_discard_trees=4)
        part.dict2model(base.model2dict())  # the documented way a fitted model travels (forest.py:161-184)
        models.append(part)
    merged = models[0]
    per_model_votes = []
    rows = frame.drop(columns=["target"])
    sample = rows.iloc[:2000]
    for mdl in models:
        per_model_votes.append(_probs_matrix(mdl.predict_proba(sample), mdl.class_vals) * len(mdl.data))
    for other in models[1:]:
        merged.mergeForest(other, by="add")
    assert len(merged.data) == 4096 and len(merged.survived_scores) == 4096
    # hard, unweighted: merged vote counts == sum of the models' vote counts (exact in fp64)
    P = _probs_matrix(merged.predict_proba(sample), merged.class_vals)
    assert np.array_equal(P * 4096, np.sum(per_model_votes, axis=0))
    # weighted soft voting over the whole frame: 8 contiguous row shards == one piece, bit for bit
    merged.setSoftVoting(True)
    merged.setWeightedTrees(True)
    whole = merged.predict_proba(rows)
    assert len(whole) == 1_000_000
    bounds = np.linspace(0, len(rows), 9).astype(int)
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        assert probs_equal(merged.predict_proba(rows.iloc[lo:hi]), whole[lo:hi])
    want = orc.forest_probs([t.data for t in merged.data], list(merged.survived_scores), merged.class_vals, rows.iloc[:24], soft=True, weighted=True)
    assert probs_equal(whole[:24], want)


def test_wide_frame_1500_features_matches_oracle():
    """Rows too wide for the 128-row shared-memory tiles (1500 features x 2-byte codes = 3 KB per row):
    the frame encoder shrinks its tile, the vote kernel reads rows from global memory, the slot gather
    and the grower take candidates of up to 1500 features.  Forest and votes == the oracle's."""
    from oracle import rfmc_oracle as orc
    from random_forest_mc_b200.model import RandomForestMC
    from tests.helpers import tree_diff

    logging.disable(logging.WARNING)
    r = np.random.default_rng(11)
    n, F = 3000, 1500
    cols = {f"x_{j}": r.integers(-20, 20, size=n).astype(np.int8) for j in range(F - 3)}
    cols["wide_0"] = np.round(r.normal(size=n), 3)            # > 255 distinct values -> uint16 codes for the whole matrix
    cols["cat_0"] = r.choice(np.array(["a", "b", "c"], dtype=object), size=n)
    cols["f32_0"] = np.round(r.normal(size=n), 1).astype(np.float32)
    frame = pd.DataFrame(cols)
    s = frame["x_0"] + frame["x_1"] + 10 * frame["wide_0"] + r.normal(size=n) * 5
    frame["target"] = pd.qcut(s, 3, labels=["lo", "mid", "hi"]).astype(str)
    kw = dict(n_trees=3, target_col="target", max_discard_trees=2, batch_train_pclass=40, batch_val_pclass=15)
    ds = orc.ingest(frame.copy(), target_col="target", batch_train_pclass=40, batch_val_pclass=15, max_depth=None, min_samples_split=1)
    np.random.seed(11)
    random.seed(11)
    want_trees, want_scores, _ = orc.fit(ds, 3, max_discard_trees=2)
    m = RandomForestMC(**kw)
    np.random.seed(11)
    random.seed(11)
    m.fit(frame.copy(), disable_progress_bar=True)
    assert m.encoded.width == 2 and m.encoded.stride * 2 > 1600
    assert [float(x) for x in m.survived_scores] == [float(x) for x in want_scores]
    for a, b in zip(m.data, want_trees):
        assert tree_diff(a.data, b) is None
    pred = frame.drop(columns=["target"]).head(300).copy()
    pred["wide_0"] = pred["wide_0"].where(r.random(300) > 0.2)
    for fr in (pred, pred.drop(columns=["x_5", "cat_0"])):
        for soft, weighted in ((False, False), (True, True)):
            m.setSoftVoting(soft)
            m.setWeightedTrees(weighted)
            assert probs_equal(m.testForestProbs(fr), orc.forest_probs(want_trees, want_scores, ds.class_vals, fr, soft, weighted))
