"""CPU, build container only: the oracle against the REAL reference run live (skipped where
/root/reference does not exist, e.g. on the GPU box).  Same randomized frames and hyper-parameters
as tests/test_gpu_fuzz.py, so the engine == oracle (GPU fuzz) and oracle == reference (here) chains
cover the same cases: forests dict for dict (value types, key order), survived scores, both RNG end
states, and votes in all four modes on frames with NaN cells and a dropped column."""
import logging
import os
import random
import sys

import numpy as np
import pytest

from oracle import rfmc_oracle as orc
from tests.helpers import probs_equal, tree_diff
from tests.test_gpu_fuzz import case_config, oracle_ingest

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference is only present in the build container")


def _reference_class():
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from random_forest_mc.model import RandomForestMC

    return RandomForestMC


CASES = [(s, False) for s in list(range(200, 230)) + [1001, 1130, 1189, 1304]] + [(s, True) for s in range(40)]


@pytest.mark.parametrize("seed,wide", CASES)
def test_oracle_equals_live_reference(seed, wide):
    logging.disable(logging.WARNING)
    Ref = _reference_class()
    frame, kw = case_config(seed, wide)
    ref = Ref(**kw)
    np.random.seed(seed)
    random.seed(seed)
    try:
        ref.fit(frame.copy(), disable_progress_bar=True)
        ref_failed = None
    except Exception as e:  # every candidate of a slot scored 0.0 (model.py:345), empty hold-out set (:306), ...
        ref_failed = type(e)
    ref_state = (np.random.get_state()[1].copy(), np.random.get_state()[2], random.getstate())

    ds = oracle_ingest(frame.copy(), kw)
    np.random.seed(seed)
    random.seed(seed)
    try:
        trees, scores, _ = orc.fit(ds, kw["n_trees"], th_start=kw["th_start"], delta_th=kw["delta_th"],
                                   max_discard_trees=kw["max_discard_trees"], get_best_tree=kw["get_best_tree"])
        failed = AttributeError if any(t is None for t in trees) else None
    except Exception as e:
        failed = type(e)
    assert failed == ref_failed
    if failed is IndexError:  # dies in the first validationTree: the streams stop at the same place
        assert np.array_equal(np.random.get_state()[1], ref_state[0]) and np.random.get_state()[2] == ref_state[1]
        assert random.getstate() == ref_state[2]
    if failed:
        return
    assert np.array_equal(np.random.get_state()[1], ref_state[0]) and np.random.get_state()[2] == ref_state[1]
    assert random.getstate() == ref_state[2]
    assert len(trees) == len(ref.data)
    assert [repr(float(s)) for s in scores] == [repr(float(s)) for s in ref.survived_scores]
    for a, b in zip(trees, ref.data):
        assert tree_diff(a, b.data) is None
    assert list(ds.class_vals) == list(ref.class_vals)

    pred = frame.drop(columns=["target"]).head(40).copy()
    numeric = [c for c in pred.columns if pred[c].dtype.kind in "fiu"]
    if numeric:
        c0 = numeric[0]
        pred[c0] = pred[c0].astype("float64" if pred[c0].dtype != np.float32 else "float32")
        pred.loc[pred.index[::4], c0] = np.nan
    frames = [pred]
    if len(pred.columns) > 1:
        frames.append(pred.drop(columns=[pred.columns[-1]]))
    for fr in frames:
        for soft in (False, True):
            for weighted in (False, True):
                if weighted and not all(s == s and s > 0 for s in scores):
                    continue
                ref.setSoftVoting(soft)
                ref.setWeightedTrees(weighted)
                assert probs_equal(orc.forest_probs(trees, scores, ds.class_vals, fr, soft, weighted), ref.testForestProbs(fr))
                assert orc.forest_labels(trees, scores, ds.class_vals, fr, soft, weighted) == ref.testForest(fr)
