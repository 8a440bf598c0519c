"""CPU, build container only (skipped where /root/reference is absent): degenerate datasets and constructor arguments
on the drop-in API and on the REAL reference -- same outcome (forest, attributes, votes, caller's frame) or the same
exception type, with both RNG streams in the same state afterwards."""
import logging
import os
import random
import sys

import numpy as np
import pandas as pd
import pytest

from tests import fake_engine
from tests.helpers import probs_equal, tree_diff

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference is only present in the build container")

_g = np.random.default_rng(0)
_n = 60
BASE = pd.DataFrame({"x_1": _g.normal(size=_n).round(1), "y_2": _g.integers(0, 5, size=_n), "c_3": _g.choice(list("ab"), size=_n),
                     "target": _g.choice(["u", "v", "w"], size=_n)})
CASES = {
    "normal": (BASE, {}),
    "single class": (BASE.assign(target="u"), {}),
    "class with one row": (pd.concat([BASE[BASE.target != "w"], BASE[BASE.target == "w"].head(1)]), {}),
    "all rows have nan": (BASE.assign(x_1=np.nan), {}),
    "only the target column": (BASE[["target"]], {}),
    "one feature, default min_feature": (BASE[["x_1", "target"]], {}),
    "missing target column": (BASE.drop(columns=["target"]), {}),
    "int target": (BASE.assign(target=_g.integers(0, 3, size=_n)), {}),
    "bool target": (BASE.assign(target=_g.random(_n) < 0.5), {}),
    "float target with nan": (BASE.assign(target=np.where(_g.random(_n) < 0.2, np.nan, 1.0)), {}),
    "one tree, one candidate": (BASE, dict(n_trees=1, max_discard_trees=1)),
    "max_depth 1": (BASE, dict(max_depth=1)),
    "max_depth 0": (BASE, dict(max_depth=0)),
    "min_samples_split larger than the data": (BASE, dict(min_samples_split=1000)),
    "min_feature = max_feature = 1": (BASE, dict(min_feature=1, max_feature=1)),
    "min_feature > number of features": (BASE, dict(min_feature=7, max_feature=9)),
    "max_feature < min_feature": (BASE, dict(min_feature=3, max_feature=2)),
    "batch_val_pclass 0": (BASE, dict(batch_val_pclass=0)),
    "batch_train_pclass 0": (BASE, dict(batch_train_pclass=0)),
    "th_start 0": (BASE, dict(th_start=0.0)),
    "no best tree, threshold starts above 1": (BASE, dict(get_best_tree=False, th_start=2.0, delta_th=0.5)),
    "temporal features": (BASE, dict(temporal_features=True)),
    "temporal features, unorderable names": (BASE.rename(columns={"x_1": "x"}), dict(temporal_features=True)),
    "duplicated rows": (pd.concat([BASE.head(5)] * 12), {}),
    "empty frame": (BASE.iloc[:0], {}),
    # column names the reference's tree2feats regex (model.py:40, 413-417) treats in its own way
    "int column names": (BASE.rename(columns={"x_1": 1, "y_2": 2, "c_3": 3}), {}),
    "names with quotes": (BASE.rename(columns={"x_1": "it's", "y_2": 'say "hi"'}), {}),
    "names with blanks and a colon": (BASE.rename(columns={"x_1": "a b", "y_2": "k: v"}), {}),
    "feature named split / depth": (BASE.rename(columns={"y_2": "split", "c_3": "depth"}), {}),
    "feature named like a class label": (BASE.rename(columns={"y_2": "u"}), {}),
    "unicode names and labels": (BASE.rename(columns={"x_1": "größe", "c_3": "名前"}).assign(target=_g.choice(["ä", "β", "字"], size=_n)), {}),
    "string index": (BASE.set_index(np.array([f"r{i}" for i in range(_n)])), {}),
    "datetime feature": (BASE.assign(d_4=pd.to_datetime("2020-01-01") + pd.to_timedelta(_g.integers(0, 50, size=_n), unit="D")), {}),
    "category dtype features": (BASE.astype({"c_3": "category", "y_2": "category"}), {}),
    "nullable Int64 / Float64 features": (BASE.astype({"y_2": "Int64", "x_1": "Float64"}), {}),
    "object-typed integers": (BASE.astype({"y_2": object}), {}),
    "bytes feature": (BASE.assign(b_6=[b"a", b"bb"] * (_n // 2)), {}),
}


def _call(fn, *a, **k):
    try:
        return ("ok", fn(*a, **k))
    except Exception as e:
        return ("raised", type(e).__name__)


@pytest.mark.parametrize("name", list(CASES))
def test_degenerate_case(name):
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from random_forest_mc.model import RandomForestMC as Ref

    from random_forest_mc_b200.model import RandomForestMC as Mine

    mp = pytest.MonkeyPatch()
    fake_engine.install(mp)
    try:
        logging.disable(logging.WARNING)
        frame, extra = CASES[name]
        kw = dict(n_trees=3, target_col="target", max_discard_trees=3, batch_train_pclass=5, batch_val_pclass=4)
        kw.update(extra)
        out = []
        for cls in (Mine, Ref):
            m = cls(**kw)
            np.random.seed(1)
            random.seed(1)
            f = frame.copy()
            r = _call(m.fit, f, disable_progress_bar=True)
            st = np.random.get_state()
            out.append((r if r[0] == "raised" else ("ok",), m, (st[1].tobytes(), int(st[2]), random.getstate()), f))
        (ra, a, sa, fa), (rb, b, sb, fb) = out
        assert ra == rb
        assert sa == sb, "RNG streams differ"
        assert fa.equals(fb), "the caller's frame was mutated differently"
        if ra[0] == "ok":
            assert len(a.data) == len(b.data) and all(tree_diff(x.data, y.data) is None for x, y in zip(a.data, b.data))
            assert [repr(float(s)) for s in a.survived_scores] == [repr(float(s)) for s in b.survived_scores]
            assert a.class_vals == b.class_vals and a.feature_cols == b.feature_cols and a._N == b._N
            assert [sorted(map(str, t.used_features)) for t in a.data] == [sorted(map(str, t.used_features)) for t in b.data]
            assert a.trees2depths == b.trees2depths and a.featImportance() == b.featImportance()
            pf = frame.drop(columns=["target"], errors="ignore").head(10)
            pa, pb = _call(a.testForestProbs, pf), _call(b.testForestProbs, pf)
            assert pa[0] == pb[0] and (pa[0] != "ok" or probs_equal(pa[1], pb[1]))
    finally:
        mp.undo()
