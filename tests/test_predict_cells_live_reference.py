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
