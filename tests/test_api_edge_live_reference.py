"""CPU, build container only (skipped where /root/reference is absent): the drop-in API on odd inputs
against the REAL reference run live, with the oracle-backed fake engine under the host logic --
empty / one-row frames, an extra unknown column, unseen categories, out-of-range values, permuted
columns, Series and dict rows, single-tree calls, and wrong argument types (same exception type)."""
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


@pytest.fixture(scope="module")
def pair():
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from random_forest_mc.model import RandomForestMC as Ref

    from random_forest_mc_b200.model import RandomForestMC as Mine

    mp = pytest.MonkeyPatch()
    fake_engine.install(mp)
    logging.disable(logging.WARNING)
    g = load_golden("mixed_default")
    kw = dict(g["kwargs"])
    kw["n_trees"] = 4

    def fit(cls):
        m = cls(**kw)
        np.random.seed(3)
        random.seed(3)
        m.fit(g["frame"].copy(), disable_progress_bar=True)
        return m

    mine, ref = fit(Mine), fit(Ref)
    yield mine, ref, g["frame"].drop(columns=[kw["target_col"]])
    mp.undo()


def _call(fn, *args):
    try:
        return fn(*args)
    except Exception as e:  # compared by type
        return ("raised", type(e).__name__)


def _frames(fr):
    out = {"empty": fr.iloc[:0], "one": fr.iloc[:1], "permuted": fr.head(20)[list(fr.columns[::-1])]}
    extra = fr.head(20).copy()
    extra["zzz_new"] = 1.0
    out["extra_column"] = extra
    unseen = fr.head(20).copy()
    unseen.loc[unseen.index[:5], "cat"] = "NEVER_SEEN"
    out["unseen_category"] = unseen
    num = [c for c in fr.columns if fr[c].dtype.kind in "fiu"][0]
    big = fr.head(20).copy()
    big[num] = big[num].astype(float) * 1e9
    out["out_of_range"] = big
    return out


@pytest.mark.parametrize("soft", [False, True])
@pytest.mark.parametrize("weighted", [False, True])
def test_odd_frames(pair, soft, weighted):
    mine, ref, fr = pair
    for m in (mine, ref):
        m.setSoftVoting(soft)
        m.setWeightedTrees(weighted)
    for name, f in _frames(fr).items():
        pa, pb = _call(mine.testForestProbs, f), _call(ref.testForestProbs, f)
        if isinstance(pa, tuple) or isinstance(pb, tuple):
            assert pa == pb, name
        else:
            assert probs_equal(pa, pb), name
        assert _call(mine.testForest, f) == _call(ref.testForest, f), name
        assert _call(mine.predict, f) == _call(ref.predict, f), name


def test_rows_trees_and_bad_arguments(pair):
    mine, ref, fr = pair
    row = fr.iloc[3]
    for soft in (False, True):
        for m in (mine, ref):
            m.setSoftVoting(soft)
        assert mine.predict(row) == ref.predict(row)
        assert mine.predict_proba(row) == ref.predict_proba(row)
        assert mine.useForest(row) == ref.useForest(row)
        assert _call(mine.useForest, row.to_dict()) == _call(ref.useForest, row.to_dict())
    assert [t(row) for t in mine.data] == [t(row) for t in ref.data]
    for bad in (5, "x", [1, 2], None, fr.to_numpy()):
        for name in ("predict", "predict_proba"):
            ra, rb = _call(getattr(mine, name), bad), _call(getattr(ref, name), bad)
            assert isinstance(ra, tuple) and ra == rb, (name, type(bad).__name__, ra, rb)
