"""CPU, build container only (skipped where /root/reference is absent): predict-time frames with awkward cells against
the REAL reference run live -- every column dtype the frame may arrive in (float16...float64, object-typed numbers,
+-inf, Int64, category, all-float32 frames) and cells of numeric features that are not numbers (None, pd.NA, strings):
the reference raises TypeError exactly when some tree's walk reaches a node that tests such a cell (tree.py:177-180),
also through the both-branches evaluation of a column that is missing from the frame (tree.py:166-175)."""
import logging
import os
import random
import sys

import numpy as np
import pandas as pd
import pytest

from tests import fake_engine
from tests.helpers import load_golden, probs_equal

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference is only present in the build container")


def _call(fn, *a, **k):
    try:
        return ("ok", fn(*a, **k))
    except Exception as e:
        return ("raised", "TypeError" if isinstance(e, TypeError) else type(e).__name__)


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
    kw["n_trees"] = 5
    out = []
    for cls in (Mine, Ref):
        m = cls(**kw)
        np.random.seed(2)
        random.seed(2)
        m.fit(g["frame"].copy(), disable_progress_bar=True)
        out.append(m)
    yield out[0], out[1], g["frame"].drop(columns=[kw["target_col"]]).head(40).copy()
    mp.undo()


def _same(a, b, frame):
    pa, pb = _call(a.testForestProbs, frame), _call(b.testForestProbs, frame)
    return pa[0] == pb[0] and (pa == pb if pa[0] == "raised" else probs_equal(pa[1], pb[1]))


def test_column_dtypes(pair):
    a, b, fr = pair
    num = [c for c in fr.columns if fr[c].dtype.kind in "fiu"]
    frames = {"numeric only, float32": fr[num].astype("float32"), "numeric only": fr[num], "float16": fr.astype({num[0]: "float16"})}
    for c in num:
        frames[f"{c} float64"] = fr.astype({c: "float64"})
        frames[f"{c} float32"] = fr.astype({c: "float32"})
        frames[f"{c} object"] = fr.astype({c: object})
        frames[f"{c} str"] = fr.astype({c: str})
        frames[f"{c} inf"] = fr.astype({c: "float64"}).assign(**{c: np.inf})
        frames[f"{c} bool"] = fr.assign(**{c: fr[c] > 0})
        frames[f"{c} category"] = fr.astype({c: "category"})
        if fr[c].dtype.kind in "iu":
            frames[f"{c} Int64"] = fr.astype({c: "Int64"})
    frames["cat as category"] = fr.astype({"cat": "category"})
    frames["cat as float"] = fr.assign(cat=1.5)
    for soft in (False, True):
        for m in (a, b):
            m.setSoftVoting(soft)
        for name, f in frames.items():
            assert _same(a, b, f), (name, soft)


@pytest.mark.parametrize("trial", range(120))
def test_non_numeric_cells_raise_only_when_reached(pair, trial):
    a, b, fr = pair
    rg = np.random.default_rng(trial)
    f = fr.copy().astype(object) if trial % 3 == 0 else fr.copy()
    c = rg.choice([c for c in fr.columns if fr[c].dtype.kind in "fiu"])
    r = int(rg.integers(0, len(f)))
    kind = trial % 4
    if kind == 0:
        f = f.astype({c: object})
        f.iloc[r, f.columns.get_loc(c)] = None
    elif kind == 1:
        f = f.astype({c: object})
        f.iloc[r, f.columns.get_loc(c)] = "oops"
    elif kind == 2 and fr[c].dtype.kind in "iu":
        f = f.astype({c: "Int64"})
        f.iloc[r, f.columns.get_loc(c)] = pd.NA
    else:
        f = f.astype({c: "category"})
    if trial % 5 == 0:  # plus a column missing from the frame: both subtrees are evaluated below its nodes
        f = f.drop(columns=[rg.choice([x for x in f.columns if x != c])])
    assert _same(a, b, f.iloc[[r]])
    assert _same(a, b, f)
    assert _call(a.useForest, f.iloc[r]) == _call(b.useForest, f.iloc[r])


def test_float32_cells_meet_python_float_split_values_in_float32():
    """Trained on float64 columns, voted on float32 frames / rows / dict rows / single-tree calls, fitted and
    JSON-reloaded: a np.float32 cell is compared with a Python-float split value in float32 (NEP 50), so a cell equal
    to the float32 rounding of a split value goes to '>=' even when the rounding went down."""
    import json

    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from random_forest_mc.model import RandomForestMC as Ref
    from random_forest_mc.utils import json_encoder

    from random_forest_mc_b200.model import RandomForestMC as Mine

    mp = pytest.MonkeyPatch()
    fake_engine.install(mp)
    try:
        logging.disable(logging.WARNING)
        g = np.random.default_rng(0)
        n = 300
        frame = pd.DataFrame({"a_1": g.normal(size=n).round(2), "b_2": g.normal(size=n).round(1), "c_3": g.normal(size=n).round(3),
                              "d_4": g.integers(-300, 300, size=n).astype(np.int16), "e_5": g.integers(0, 100000, size=n).astype(np.int32)})
        frame["target"] = np.where(frame.a_1 + frame.b_2 + g.normal(size=n) * 0.3 > 0, "p", "q")
        kw = dict(n_trees=6, target_col="target", max_discard_trees=3, batch_train_pclass=40, batch_val_pclass=20)
        fitted = []
        for cls in (Mine, Ref):
            m = cls(**kw)
            np.random.seed(1)
            random.seed(1)
            m.fit(frame.copy(), disable_progress_bar=True)
            fitted.append(m)
        a, b = fitted
        a2, b2 = Mine(**kw), Ref(**kw)
        a2.dict2model(json.loads(json.dumps(a.model2dict(), default=json_encoder)))
        b2.dict2model(json.loads(json.dumps(b.model2dict(), default=json_encoder)))
        X = frame.drop(columns=["target"])
        f32 = {"a_1": "float32", "b_2": "float32", "c_3": "float32"}
        frames = {"float32 + int16 (float32 rows)": X.drop(columns=["e_5"]).astype(f32), "float32 + int32 (float64 rows)": X.astype(f32),
                  "all float32": X.astype("float32"), "two float32 columns": X[["a_1", "b_2"]].astype("float32"),
                  "float32 next to a string column (object rows)": X.astype(f32).assign(s="z")}
        differs_in_float64 = 0
        for ma, mb in ((a, b), (a2, b2)):
            for soft in (False, True):
                for m in (ma, mb):
                    m.setSoftVoting(soft)
                    m.setWeightedTrees(soft)
                for name, f in frames.items():
                    got = ma.testForestProbs(f)
                    assert probs_equal(got, mb.testForestProbs(f)), name
                    differs_in_float64 += got != ma.testForestProbs(f.astype("float64")) if "s" not in f.columns else 0
                    for i in range(0, len(f), 11):
                        row = f.iloc[i]
                        assert ma.useForest(row) == mb.useForest(row), name
                        as_dict = {k: row[k] for k in row.index}
                        assert ma.useForest(as_dict) == mb.useForest(as_dict), name
                        assert [t(row) for t in ma.data] == [t(row) for t in mb.data], name
        assert differs_in_float64 > 0  # the case is not vacuous: float32 rows do vote differently from their float64 copies
    finally:
        mp.undo()
