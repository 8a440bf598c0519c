"""Checks of the drop-in Python API against the vectors recorded from the reference.  Used twice:
with the oracle-backed fake engine on CPU (host logic) and with the real engine on the GPU."""
import hashlib
import json
import random

import numpy as np

from random_forest_mc_b200.model import RandomForestMC
from tests.helpers import load_golden, probs_equal, tree_diff


def rng_fingerprint():
    st = np.random.get_state()
    h = hashlib.sha256()
    h.update(st[1].tobytes())
    h.update(repr(st[2:]).encode())
    h.update(repr(random.getstate()).encode())
    return h.hexdigest()


def fitted_model(name, **overrides):
    g = load_golden(name)
    m = RandomForestMC(**g["kwargs"])
    for k, v in overrides.items():
        setattr(m, k, v)
    np.random.seed(g["seed"])
    random.seed(g["seed"])
    m.fit(g["frame"].copy(), disable_progress_bar=True)
    return g, m


def check_fit_matches_reference(name, **overrides):
    """Same seeds -> identical forest (tree dicts incl. types and key order), identical survived
    scores, identical number of candidates consumed, and both RNG streams left exactly where the
    reference leaves them (so the speculation/rewind is invisible)."""
    g, m = fitted_model(name, **overrides)
    assert m.class_vals == g["class_vals"]
    assert m.feature_cols == g["feature_cols"]
    assert m.numeric_cols == g["numeric_cols"]
    assert m.type_of_cols == g["type_of_cols"]
    assert m._N == g["rows_per_class"]
    assert len(m.data) == len(g["forest"])
    assert [float(s) for s in m.survived_scores] == [float(s) for s in g["survived_scores"]]
    for mine, ref in zip(m.data, g["forest"]):
        assert tree_diff(mine.data, ref) is None
    assert m.fit_stats["candidates_used"] == len(g["trace"])
    assert rng_fingerprint() == g["rng_after"]
    assert [t.md5hexdigest for t in m.data] == g["md5"] or any("-0.0" in json.dumps(t, default=str) for t in g["forest"])
    for t in m.data:
        assert t.features == g["feature_cols"]
        assert set(t.used_features) <= set(g["feature_cols"])
    return g, m


def check_predictions_match_reference(g, m):
    for (fname, soft, weighted), want in g["preds"].items():
        m.setSoftVoting(soft)
        m.setWeightedTrees(weighted)
        fr = g["frames"][fname]
        assert m.testForest(fr) == want["labels"], (fname, soft, weighted)
        assert probs_equal(m.testForestProbs(fr), want["probs"]), (fname, soft, weighted)
        assert m.predict(fr) == want["labels"]
        assert probs_equal(m.predict_proba(fr), want["probs"])
        row = fr.iloc[3]
        got = m.predict(row)
        assert list(got.keys()) == g["class_vals"]
        assert all(float(got[c]) == float(want["probs"][3][c]) for c in g["class_vals"]), (fname, soft, weighted)


def check_persistence_and_merge(g, m):
    d = json.loads(json.dumps(m.model2dict(), default=lambda o: o.item() if isinstance(o, np.generic) else None))
    m2 = RandomForestMC(target_col=g["kwargs"]["target_col"])
    m2.dict2model(d)
    assert len(m2.data) == len(m.data)
    for soft in (False, True):
        for weighted in (False, True):
            for mm in (m, m2):
                mm.setSoftVoting(soft)
                mm.setWeightedTrees(weighted)
            fr = g["frames"]["nan"]
            assert m2.testForest(fr) == m.testForest(fr)
            assert probs_equal(m2.testForestProbs(fr), m.testForestProbs(fr))
    n0 = len(m2.data)
    m2.mergeForest(m, by="add")
    assert len(m2.data) == 2 * n0 and len(m2.survived_scores) == 2 * n0
    kept = m2.drop_duplicated_trees()
    assert kept == len(m2.data) <= 2 * n0
    m3 = RandomForestMC(target_col=g["kwargs"]["target_col"])
    m3.dict2model(d)
    m3.mergeForest(m, N=3, by="score")
    assert len(m3.data) == 3
    assert m3.survived_scores == sorted(m3.survived_scores, reverse=True)


def check_predict_missing_values():
    """predictMissingValues == the frame recorded from the reference (same seed, same inputs)."""
    import gzip
    import os
    import pickle

    import pandas as pd

    from tests.helpers import GOLDEN_DIR

    with gzip.open(os.path.join(GOLDEN_DIR, "..", "golden_missing", "missing_titanic.pkl.gz"), "rb") as f:
        g = pickle.load(f)
    m = RandomForestMC(**g["kwargs"])
    np.random.seed(g["seed"])
    random.seed(g["seed"])
    m.fit(g["frame"].copy(), disable_progress_bar=True)
    assert [float(s) for s in m.survived_scores] == [float(s) for s in g["survived_scores"]]
    got_row = m.predictMissingValues(g["row"], g["dict_values"])
    got_frame = m.predictMissingValues(g["holes"], g["dict_values"])
    pd.testing.assert_frame_equal(got_row, g["by_row"], check_exact=True, check_dtype=False)
    pd.testing.assert_frame_equal(got_frame, g["by_frame"], check_exact=True, check_dtype=False)
    from random_forest_mc_b200.model import DictValuesAllFeaturesMissing, MissingValuesNotFound
    import pytest

    with pytest.raises(MissingValuesNotFound):
        m.predictMissingValues(g["frame"].loc[0], g["dict_values"])
    with pytest.raises(DictValuesAllFeaturesMissing):
        m.predictMissingValues(g["holes"], {"Not Feature": [1, 2], "Not Feature 2": [3]})


def check_pickle_and_deepcopy():
    """The reference model is a plain picklable object (its fitParallel pickles it; users joblib.dump it).  A copy
    travels without device handles: it votes at once (the forest is re-flattened) and re-encodes its dataset on the
    device when it is fitted again, continuing exactly as the original would."""
    import copy
    import pickle
    import random

    from random_forest_mc_b200.model import RandomForestMC

    g = load_golden("mixed_default")
    kw = dict(g["kwargs"])
    kw["n_trees"] = 3
    m = RandomForestMC(**kw)
    np.random.seed(1)
    random.seed(1)
    m.fit(g["frame"].copy(), disable_progress_bar=True)
    fr = g["frame"].drop(columns=[kw["target_col"]]).head(20)
    m.setSoftVoting(True)
    want = m.testForestProbs(fr)
    copies = [copy.deepcopy(m), pickle.loads(pickle.dumps(m))]
    for c in copies:
        assert c.testForestProbs(fr) == want and c.soft_voting
        assert all(tree_diff(a.data, b.data) is None for a, b in zip(c.data, m.data))
        np.random.seed(2)
        random.seed(2)
        c.fit(disable_progress_bar=True)
    np.random.seed(2)
    random.seed(2)
    m.fit(disable_progress_bar=True)
    for c in copies:
        assert len(c.data) == len(m.data) == 6
        assert all(tree_diff(a.data, b.data) is None for a, b in zip(c.data, m.data))
        assert [float(s) for s in c.survived_scores] == [float(s) for s in m.survived_scores]
