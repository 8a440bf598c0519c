"""CPU, build container only (skipped where /root/reference is absent): STATEFUL call sequences on the drop-in API and
on the REAL reference, step by step under the same seeds -- fit(frame), fit() again (appends), predict / predict_proba
on frames and rows, process_dataset again, fit, reset_forest, survivedTree, fit with n_trees = 0.  After every step:
same return value or same exception type, same model2dict JSON (used_features as sets, sign of zero split values
ignored: DESIGN.md deviation 1), same public attributes, and both RNG streams in the same state -- also after a fit that
dies (every candidate of a slot scoring 0.0, or an empty hold-out set)."""
import json
import logging
import os
import random
import sys

import numpy as np
import pytest

from tests import fake_engine
from tests.helpers import probs_equal
from tests.test_gpu_fuzz import case_config

REF_SRC = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference is only present in the build container")
ATTRS = ("min_feature", "max_feature", "_N", "n_samples", "class_vals", "feature_cols", "numeric_cols", "temporal_features")


def _snap(m, enc):
    d = m.model2dict()
    for t in d["Forest"]:
        t["used_features"] = sorted(t["used_features"])
    return json.dumps(d, default=enc).replace("-0.0,", "0.0,")


def _rng():
    st = np.random.get_state()
    return st[1].tobytes(), int(st[2]), random.getstate()


def _run(seed, fn):
    np.random.seed(seed)
    random.seed(seed)
    try:
        out = fn()
    except Exception as e:
        out = ("raised", type(e).__name__)
    return out, _rng()


@pytest.mark.parametrize("seed", range(40))
def test_call_sequences(seed):
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from random_forest_mc.model import RandomForestMC as Ref
    from random_forest_mc.utils import json_encoder

    from random_forest_mc_b200.model import RandomForestMC as Mine

    mp = pytest.MonkeyPatch()
    fake_engine.install(mp)
    try:
        logging.disable(logging.WARNING)
        frame, kw = case_config(seed, wide=bool(seed % 2))
        a, b = Mine(**kw), Ref(**kw)
        fr = frame.drop(columns=["target"]).head(30)

        def votes(m):
            return m.predict(fr), m.predict_proba(fr), m.predict(fr, prob_output=True), m.predict_proba(fr.iloc[0]), m.predict(fr.iloc[1])

        def zero(m):
            m.n_trees = 0
            return m.fit(disable_progress_bar=True)

        steps = [
            ("fit(frame)", lambda m: m.fit(frame.copy(), disable_progress_bar=True)),
            ("fit() appends", lambda m: m.fit(disable_progress_bar=True)),
            ("votes", votes),
            ("process_dataset again", lambda m: m.process_dataset(frame.copy())),
            ("fit after process_dataset", lambda m: m.fit(disable_progress_bar=True)),
            ("reset_forest", lambda m: m.reset_forest()),
            ("survivedTree", lambda m: m.survivedTree().data),
            ("fit with n_trees = 0", zero),
        ]
        for k, (name, fn) in enumerate(steps):
            (ra, sa), (rb, sb) = _run(100 + k, lambda: fn(a)), _run(100 + k, lambda: fn(b))
            if name == "votes" and not (isinstance(ra, tuple) and ra and ra[0] == "raised"):
                assert ra[0] == rb[0] and probs_equal(ra[1], rb[1]) and probs_equal(ra[2], rb[2]) and ra[3] == rb[3] and ra[4] == rb[4], name
            elif name == "survivedTree" and isinstance(ra, dict):
                from tests.helpers import tree_diff

                assert tree_diff(ra, rb) is None, name
            else:
                assert ra == rb, (name, ra, rb)
            assert sa == sb, f"{name}: RNG streams differ"
            assert _snap(a, json_encoder) == _snap(b, json_encoder), name
            assert all(getattr(a, x, None) == getattr(b, x, None) for x in ATTRS), name
            if isinstance(ra, tuple) and ra and ra[0] == "raised" and name.startswith("fit(frame)"):
                break  # nothing fitted: the later steps have no forest to work on
    finally:
        mp.undo()
