"""CPU, build container only (skipped where /root/reference is absent): persistence, merge, de-duplication and the
explainability helpers of the drop-in API against the REAL reference run live under the same seeds (oracle-backed fake
engine under the host logic).  `used_features` is a hash-ordered list in the reference (model.py:413-417), so lists of
features are compared as sets and dicts without regard to key order."""
import json
import logging
import os
import random
import sys

import numpy as np
import pytest

from tests import fake_engine
from tests.helpers import load_golden, probs_equal

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference is only present in the build container")


@pytest.fixture(scope="module", params=["mixed_default", "titanic_c1", "ties_default"])
def ctx(request):
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from random_forest_mc.model import RandomForestMC as Ref
    from random_forest_mc.utils import json_encoder

    from random_forest_mc_b200.model import RandomForestMC as Mine

    mp = pytest.MonkeyPatch()
    fake_engine.install(mp)
    logging.disable(logging.WARNING)
    g = load_golden(request.param)
    kw = dict(g["kwargs"])
    kw["n_trees"] = 6

    def fit(cls, seed=5):
        m = cls(**kw)
        np.random.seed(seed)
        random.seed(seed)
        m.fit(g["frame"].copy(), disable_progress_bar=True)
        return m

    yield dict(fit=fit, Mine=Mine, Ref=Ref, kw=kw, enc=json_encoder, frame=g["frame"].drop(columns=[kw["target_col"]]))
    mp.undo()


def _normalised(d):
    for t in d["Forest"]:
        t["used_features"] = sorted(t["used_features"])
    return d


def test_model2dict_json_md5_and_cross_loading(ctx):
    a, b = ctx["fit"](ctx["Mine"]), ctx["fit"](ctx["Ref"])
    ja = json.dumps(_normalised(a.model2dict()), default=ctx["enc"])
    jb = json.dumps(_normalised(b.model2dict()), default=ctx["enc"])
    assert ja == jb  # same wire format, byte for byte
    assert [t.md5hexdigest for t in a.data] == [t.md5hexdigest for t in b.data]
    assert repr(a) == repr(b) and a.Forest_size == b.Forest_size and a.trees2depths == b.trees2depths
    mine_from_ref, ref_from_mine = ctx["Mine"](**ctx["kw"]), ctx["Ref"](**ctx["kw"])
    mine_from_ref.dict2model(json.loads(jb))
    ref_from_mine.dict2model(json.loads(ja))
    fr = ctx["frame"].head(60)
    for soft in (False, True):
        for weighted in (False, True):
            for m in (mine_from_ref, ref_from_mine):
                m.setSoftVoting(soft)
                m.setWeightedTrees(weighted)
            assert probs_equal(mine_from_ref.testForestProbs(fr), ref_from_mine.testForestProbs(fr))
            assert mine_from_ref.testForest(fr) == ref_from_mine.testForest(fr)
    assert (a == mine_from_ref) == (b == ref_from_mine)


def test_merge_and_drop_duplicates(ctx):
    fit, Mine, Ref = ctx["fit"], ctx["Mine"], ctx["Ref"]
    for by, n in (("add", -1), ("score", 4), ("random", 5)):
        a, b = fit(Mine, 7), fit(Ref, 7)
        random.seed(99)
        a.mergeForest(fit(Mine, 8), n, by)
        random.seed(99)
        b.mergeForest(fit(Ref, 8), n, by)
        assert [t.md5hexdigest for t in a.data] == [t.md5hexdigest for t in b.data], by
        assert [float(s) for s in a.survived_scores] == [float(s) for s in b.survived_scores], by
    a, b = fit(Mine, 7), fit(Ref, 7)
    a.mergeForest(fit(Mine, 7))
    b.mergeForest(fit(Ref, 7))
    assert a.drop_duplicated_trees() == b.drop_duplicated_trees() and len(a.data) == len(b.data)


def _call(fn, *args, **kw):
    try:
        return fn(*args, **kw)
    except Exception as e:
        return ("raised", type(e).__name__)


def _equal(x, y):
    if hasattr(x, "equals"):
        return x.equals(y)
    if isinstance(x, tuple) and isinstance(y, tuple) and len(x) == 2 and isinstance(x[1], list):
        return x[0] == y[0] and x[1] == y[1]
    return x == y  # dicts: equal whatever the key order


def test_explainability_helpers(ctx):
    a, b = ctx["fit"](ctx["Mine"]), ctx["fit"](ctx["Ref"])
    for name in ("featCount", "featImportance", "featScoreMean", "featCorrDataFrame"):
        assert _equal(_call(getattr(a, name)), _call(getattr(b, name))), name
    assert a.featPairImportance(disable_progress_bar=True) == b.featPairImportance(disable_progress_bar=True)
    row = ctx["frame"].iloc[2]
    for Class in a.class_vals:
        assert [t.md5hexdigest for t in a.sampleClass2trees(row, Class)] == [t.md5hexdigest for t in b.sampleClass2trees(row, Class)]
        for name in ("sampleClassFeatCount", "sampleClassFeatImportance", "sampleClassFeatScoreMean",
                     "sampleClassFeatPairImportance", "sampleClassFeatCorrDataFrame"):
            assert _equal(_call(getattr(a, name), row, Class), _call(getattr(b, name), row, Class)), (name, Class)


def test_scores_out_of_step_with_trees(ctx):
    """dict2model aliases the dict's score list (forest.py:177-184), so scores and trees can get out of step; weighted
    votes then pair them by zip() but divide by fsum of ALL scores (forest.py:208-212, 224-231)."""
    a, b = ctx["fit"](ctx["Mine"]), ctx["fit"](ctx["Ref"])
    fr = ctx["frame"].head(40)
    for change in ("one score too many", "one score too few"):
        for m in (a, b):
            m.survived_scores = [float(s) for s in m.survived_scores]
            if change == "one score too many":
                m.survived_scores.append(0.123)
            else:
                del m.survived_scores[-2:]
        for soft in (False, True):
            for weighted in (False, True):
                for m in (a, b):
                    m.setSoftVoting(soft)
                    m.setWeightedTrees(weighted)
                assert probs_equal(a.testForestProbs(fr), b.testForestProbs(fr)), (change, soft, weighted)
                assert a.useForest(fr.iloc[1]) == b.useForest(fr.iloc[1]), (change, soft, weighted)


def test_pathological_scores_and_an_emptied_forest(ctx):
    """Weighted votes with score lists a user can produce by hand: all zero / zero sum (ZeroDivisionError), NaN, negative,
    integers, no scores at all, denormals, 1e308 (OverflowError from fsum); and a forest emptied of its trees but not
    of its scores (weighted modes return zeros, unweighted ones ZeroDivisionError)."""
    fr = ctx["frame"].head(20)
    cases = {"all zero": [0.0] * 6, "zero sum": [0.5, -0.5, 0.25, -0.25, 1.0, -1.0], "nan": [0.5, float("nan"), 0.5, 0.5, 0.5, 0.5],
             "negative": [-0.5, -0.25, -1.0, -0.1, -0.2, -0.3], "numpy zeros": [np.float64(0.0)] * 6, "integers": [1, 2, 3, 4, 5, 6],
             "no scores": [], "denormal": [1e-320] * 6, "huge": [1e308] * 6}
    a, b = ctx["fit"](ctx["Mine"]), ctx["fit"](ctx["Ref"])

    def same(x, y):
        if x[0] == "raised" or y[0] == "raised":
            return x == y
        return all(all(p[k] == q[k] or (p[k] != p[k] and q[k] != q[k]) for k in p) and list(p) == list(q) for p, q in zip(x[1], y[1]))

    def call(fn, *args):
        try:
            return ("ok", fn(*args))
        except Exception as e:
            return ("raised", type(e).__name__)

    for name, scores in cases.items():
        a.survived_scores, b.survived_scores = list(scores), list(scores)
        for soft in (False, True):
            for m in (a, b):
                m.setSoftVoting(soft)
                m.setWeightedTrees(True)
            assert same(call(a.testForestProbs, fr), call(b.testForestProbs, fr)), (name, soft)
            if name != "nan":  # the label of an all-NaN vote depends on sorted()'s handling of NaN
                assert call(a.testForest, fr) == call(b.testForest, fr), (name, soft)
    a.survived_scores, b.survived_scores = [0.5] * 6, [0.5] * 6
    a.data, b.data = [], []
    for soft in (False, True):
        for weighted in (False, True):
            for m in (a, b):
                m.setSoftVoting(soft)
                m.setWeightedTrees(weighted)
            assert same(call(a.testForestProbs, fr), call(b.testForestProbs, fr)), (soft, weighted)
            assert call(a.testForest, fr) == call(b.testForest, fr), (soft, weighted)
            assert call(a.useForest, fr.iloc[0]) == call(b.useForest, fr.iloc[0]), (soft, weighted)
