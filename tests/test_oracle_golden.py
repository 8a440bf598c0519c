"""Pins oracle/rfmc_oracle.py against vectors recorded from the real reference
(tests/golden/make_golden.py).  CPU only."""
import random

import numpy as np
import pytest

from oracle import rfmc_oracle as orc
from tests.helpers import golden_names, load_golden, probs_equal, tree_diff

POLICY_KEYS = ("th_start", "delta_th", "max_discard_trees", "get_best_tree")


def _ingest(g):
    kw = g["kwargs"]
    ds = orc.ingest(
        g["frame"].copy(),
        target_col=kw["target_col"],
        batch_train_pclass=kw.get("batch_train_pclass", 10),
        batch_val_pclass=kw.get("batch_val_pclass", 10),
        min_feature=kw.get("min_feature"),
        max_feature=kw.get("max_feature"),
        temporal_features=kw.get("temporal_features", False),
        max_depth=kw.get("max_depth"),
        min_samples_split=kw.get("min_samples_split", 1),
    )
    return ds


@pytest.mark.parametrize("name", golden_names())
def test_ingest_matches_reference(name):
    g = load_golden(name)
    ds = _ingest(g)
    assert ds.class_vals == g["class_vals"]
    assert ds.feature_cols == g["feature_cols"]
    assert ds.numeric_cols == g["numeric_cols"]
    assert ds.type_of_cols == g["type_of_cols"]
    assert ds.rows_per_class == g["rows_per_class"]
    assert (ds.min_feature, ds.max_feature) == (g["min_feature"], g["max_feature"])
    assert ds.temporal_features == g["temporal_features"]
    assert ds.n_samples == g["n_samples"]


@pytest.mark.parametrize("name", golden_names())
def test_fit_replay_matches_reference(name):
    """Same seeds -> same draws, same candidate trees (types + key order), same scores,
    same kept/discarded decisions."""
    g = load_golden(name)
    ds = _ingest(g)
    kw = g["kwargs"]
    policy = {k: kw[k] for k in POLICY_KEYS if k in kw}
    np.random.seed(g["seed"])
    random.seed(g["seed"])
    trace = []
    trees, scores, grown = orc.fit(ds, kw["n_trees"], trace=trace, **policy)
    assert grown == len(g["trace"])
    for mine, ref in zip(trace, g["trace"]):
        assert np.array_equal(mine["idx_T"], ref["idx_T"])
        assert np.array_equal(mine["idx_V"], ref["idx_V"])
        assert mine["F"] == ref["F"]
        assert tree_diff(mine["tree"], ref["tree"]) is None
        assert float(mine["th_val"]) == float(ref["th_val"])
    assert [float(s) for s in scores] == [float(s) for s in g["survived_scores"]]
    for a, b in zip(trees, g["forest"]):
        assert tree_diff(a, b) is None


@pytest.mark.parametrize("name", golden_names())
def test_grow_and_score_are_pure_functions_of_the_draws(name):
    """plantTree / validationTree depend only on (data, idx_T, idx_V, F): replay each recorded
    candidate without touching any RNG."""
    g = load_golden(name)
    ds = _ingest(g)
    for ref in g["trace"]:
        t = orc.grow(ds, ref["idx_T"], list(ref["F"]))
        assert tree_diff(t, ref["tree"]) is None
        assert float(orc.score_candidate(ds, t, ref["idx_V"])) == float(ref["th_val"])


@pytest.mark.parametrize("name", golden_names())
def test_voting_matches_reference(name):
    g = load_golden(name)
    for (fname, soft, weighted), want in g["preds"].items():
        fr = g["frames"][fname]
        probs = orc.forest_probs(g["forest"], g["survived_scores"], g["class_vals"], fr, soft, weighted)
        labels = orc.forest_labels(g["forest"], g["survived_scores"], g["class_vals"], fr, soft, weighted)
        assert probs_equal(probs, want["probs"]), (fname, soft, weighted)
        assert labels == want["labels"], (fname, soft, weighted)
