"""CPU, build container only (skipped where /root/reference is absent): the remaining public seams against the REAL
reference run live -- split_train_val / sampleFeats / plantTree / validationTree called directly, DecisionTreeMC calls
(Series and dict rows), depths, str/repr, tree2dict keys, the ordering operators incl. wrong operand types, sorted(),
and testForestParallel / testForestProbsParallel (the reference runs its process pool)."""
import logging
import os
import random
import sys

import numpy as np
import pytest

from tests import fake_engine
from tests.helpers import probs_equal, tree_diff
from tests.test_gpu_fuzz import case_config

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference is only present in the build container")


def _call(fn, *a, **k):
    try:
        return fn(*a, **k)
    except Exception as e:
        return ("raised", type(e).__name__)


@pytest.mark.parametrize("seed", range(12))
def test_single_candidate_api_tree_objects_and_parallel_votes(seed):
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from random_forest_mc.model import RandomForestMC as Ref

    from random_forest_mc_b200.model import RandomForestMC as Mine

    mp = pytest.MonkeyPatch()
    fake_engine.install(mp)
    try:
        logging.disable(logging.WARNING)
        frame, kw = case_config(seed, wide=bool(seed % 2))
        a, b = Mine(**kw), Ref(**kw)
        for m in (a, b):
            np.random.seed(seed)
            random.seed(seed)
            _call(m.fit, frame.copy(), disable_progress_bar=True)
        if not b.data:
            assert not a.data
            return
        draws = []
        for m in (a, b):
            np.random.seed(3)
            random.seed(3)
            draws.append((_call(m.split_train_val), _call(m.sampleFeats, m.feature_cols[:])))
        (sa, fa), (sb, fb) = draws
        assert np.array_equal(sa[0], sb[0]) and np.array_equal(sa[1], sb[1]) and fa == fb
        ta, tb = _call(a.plantTree, sa[0], fa[:]), _call(b.plantTree, sb[0], fb[:])
        assert tree_diff(ta.data, tb.data) is None
        va, vb = _call(a.validationTree, ta, sa[1]), _call(b.validationTree, tb, sb[1])
        assert va == vb or (va != va and vb != vb)
        row = frame.drop(columns=["target"]).iloc[4]
        assert ta(row) == tb(row) and ta(row.to_dict()) == tb(row.to_dict())
        assert ta.depths == tb.depths and repr(ta) == repr(tb)
        assert str(ta).replace("-0.0", "0.0") == str(tb).replace("-0.0", "0.0")
        assert set(ta.tree2dict()) == set(tb.tree2dict())
        if len(a.data) >= 2:
            for op in ("__gt__", "__ge__", "__lt__", "__le__", "__eq__"):
                assert _call(getattr(a.data[0], op), a.data[1]) == _call(getattr(b.data[0], op), b.data[1]), op
                assert _call(getattr(a.data[0], op), 5) == _call(getattr(b.data[0], op), 5), op
            assert [float(t.survived_score) for t in sorted(a.data)] == [float(t.survived_score) for t in sorted(b.data)]
        fr = frame.drop(columns=["target"]).head(12)
        assert _call(a.testForestParallel, fr, max_workers=2) == _call(b.testForestParallel, fr, max_workers=2)
        assert probs_equal(a.testForestProbsParallel(fr, max_workers=2), b.testForestProbsParallel(fr, max_workers=2))
    finally:
        mp.undo()


def test_hand_built_tree_called_on_awkward_rows():
    """A DecisionTreeMC built directly from a nested dict (as the reference's own unit tests do), called on Series and
    dict rows: equal values, missing features (both subtrees evaluated, first leaf renormalised), NaN / None cells, a
    string in a numeric feature (TypeError only when reached), a float32 cell, bools, an empty row."""
    import pandas as pd

    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from random_forest_mc.tree import DecisionTreeMC as Ref

    from random_forest_mc_b200.tree import DecisionTreeMC as Mine

    mp = pytest.MonkeyPatch()
    fake_engine.install(mp)
    try:
        leaf = lambda d: {"leaf": d, "depth": "3#"}  # noqa: E731
        tree = {"feat1": {"split": {"feat_type": "numeric", "split_val": 5,
                                    ">=": {"feat2": {"split": {"feat_type": "categorical", "split_val": "A",
                                                               ">=": leaf({"0": 0.8, "1": 0.2}), "<": leaf({"0": 0.3, "1": 0.7})}}},
                                    "<": {"feat3": {"split": {"feat_type": "numeric", "split_val": 2.5,
                                                              ">=": leaf({"1": 1.0}), "<": leaf({"0": 0.5, "1": 0.5})}}}}}}
        a, b = Mine(tree, ["0", "1"]), Ref(tree, ["0", "1"])
        rows = [{"feat1": 6, "feat2": "A", "feat3": 1}, {"feat1": 5, "feat2": "B", "feat3": 1}, {"feat1": 4, "feat2": "A", "feat3": 2.5},
                {"feat1": 4, "feat2": "A", "feat3": 2.4}, {"feat2": "A", "feat3": 3}, {"feat1": 7, "feat3": 3}, {"feat1": 1},
                {"feat1": np.nan, "feat2": "A", "feat3": 9}, {"feat1": 7, "feat2": None, "feat3": 9}, {},
                {"feat1": "x", "feat2": "A", "feat3": 1}, {"feat1": 7, "feat2": "A", "feat3": "x"},
                {"feat1": 5.0, "feat2": "A", "feat3": np.float32(2.5)}, {"feat1": True, "feat2": "A", "feat3": False}]
        for r in rows:
            for form in (pd.Series(r, dtype=object), r):
                assert _call(a, form) == _call(b, form), r
        assert a.depths == b.depths and str(a) == str(b) and a.md5hexdigest == b.md5hexdigest
        assert a.tree2dict().keys() == b.tree2dict().keys()
    finally:
        mp.undo()
