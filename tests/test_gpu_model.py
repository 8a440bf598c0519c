"""GPU (B200): the drop-in API end to end through the real CUDA engine, against the vectors
recorded from the reference -- identical forests, kept/discarded decisions, scores, RNG end
states, labels and bit-identical probabilities in all four voting modes."""
import numpy as np
import pytest

from tests import model_checks
from tests.helpers import golden_names

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_fit_and_predict_match_reference(name):
    g, m = model_checks.check_fit_matches_reference(name)
    model_checks.check_predictions_match_reference(g, m)


@pytest.mark.parametrize("cap", [1, 5])
def test_fit_is_independent_of_batch_size(cap):
    model_checks.check_fit_matches_reference("mixed_nobest", max_candidates_per_launch=cap)
    model_checks.check_fit_matches_reference("mixed_feat", max_candidates_per_launch=cap)


def test_persistence_and_merge():
    g, m = model_checks.fitted_model("mixed_default")
    model_checks.check_persistence_and_merge(g, m)


def test_soa_and_dict_flattening_agree():
    """Voting straight from the grower's SoA output == voting from the materialised dicts."""
    from random_forest_mc_b200.flat import forest_tables_from_dicts, forest_tables_from_soa, encode_frame
    from random_forest_mc_b200 import engine

    g, m = model_checks.fitted_model("mixed_big")
    ta = forest_tables_from_soa([t.soa for t in m.data], m.survived_scores, m.encoded, m.class_vals, m.feature_cols)
    tb = forest_tables_from_dicts([t.data for t in m.data], m.survived_scores, m.class_vals, m.feature_cols)
    fa, fb = engine.DeviceForest(ta), engine.DeviceForest(tb)
    fr = g["frame"].drop(columns=["target"])
    for soft in (False, True):
        for weighted in (False, True):
            pa, la = fa.predict(*encode_frame(ta, fr), soft, weighted)
            pb, lb = fb.predict(*encode_frame(tb, fr), soft, weighted)
            assert np.array_equal(pa, pb) and np.array_equal(la, lb)
    fa.close(), fb.close()


def test_single_tree_call_matches_oracle():
    from oracle import rfmc_oracle as orc

    g, m = model_checks.fitted_model("titanic_c1")
    fr = g["frames"]["absent_nan"]
    row = fr.iloc[5]
    for tree, ref in zip(m.data, g["forest"]):
        want = orc.tree_output(ref, g["class_vals"], row)
        got = tree(row)
        assert list(got.keys()) == list(want.keys())
        assert all(float(got[k]) == float(want[k]) for k in want)


def test_predict_missing_values_matches_reference():
    model_checks.check_predict_missing_values()


def test_fit_parallel_host_streams_contract():
    from tests.test_host_model import test_fit_parallel_host_streams_contract as check

    check()


def test_non_numeric_cells_raise_only_when_reached():
    """A None / string cell in a numeric column makes the reference's `val > split_val` raise TypeError, but only for
    rows whose walk reaches a node testing it (tree.py:177-180); the engine finds those walks with the leaf-export
    kernel.  Checked row by row against the oracle, which raises from the same Python comparison."""
    from oracle import rfmc_oracle as orc

    g, m = model_checks.fitted_model("mixed_default")
    fr = g["frame"].drop(columns=["target"]).head(60)
    numeric = [c for c in fr.columns if fr[c].dtype.kind in "fiu"]
    trees, scores = [t.data for t in m.data], list(m.survived_scores)
    raised = voted = 0
    for k, c in enumerate(numeric * 6):
        r = (7 * k) % len(fr)
        one = fr.iloc[[r]].astype({c: object})
        one.iloc[0, one.columns.get_loc(c)] = None if k % 2 else "oops"
        if k % 5 == 0:
            one = one.drop(columns=[numeric[(k + 1) % len(numeric)]]) if numeric[(k + 1) % len(numeric)] != c else one
        try:
            want = orc.forest_probs(trees, scores, g["class_vals"], one, False, False)
        except TypeError:
            want = None
        if want is None:
            with pytest.raises(TypeError):
                m.testForestProbs(one)
            raised += 1
        else:
            got = m.testForestProbs(one)
            assert [float(got[0][c]) for c in g["class_vals"]] == [float(want[0][c]) for c in g["class_vals"]]
            voted += 1
    assert raised > 0 and voted > 0


def test_float32_frames_on_a_float64_trained_forest_match_the_oracle():
    """np.float32 cells meet Python-float split values in float32 (NEP 50): the engine flattens a float32 variant of
    the forest for frames that hand over float32 rows.  The oracle walks real pandas rows with numpy comparisons, so
    it follows the same promotion rules as the reference (pinned live in tests/test_predict_cells_live_reference.py)."""
    import logging
    import random

    import pandas as pd

    from oracle import rfmc_oracle as orc
    from random_forest_mc_b200.model import RandomForestMC
    from tests.helpers import probs_equal

    logging.disable(logging.WARNING)
    g = np.random.default_rng(0)
    n = 300
    frame = pd.DataFrame({"a_1": g.normal(size=n).round(2), "b_2": g.normal(size=n).round(1), "c_3": g.normal(size=n).round(3),
                          "d_4": g.integers(-300, 300, size=n).astype(np.int16)})
    frame["target"] = np.where(frame.a_1 + frame.b_2 + g.normal(size=n) * 0.3 > 0, "p", "q")
    m = RandomForestMC(n_trees=6, target_col="target", max_discard_trees=3, batch_train_pclass=40, batch_val_pclass=20)
    np.random.seed(1)
    random.seed(1)
    m.fit(frame.copy(), disable_progress_bar=True)
    trees, scores = [t.data for t in m.data], list(m.survived_scores)
    X = frame.drop(columns=["target"])
    f32 = X.astype({"a_1": "float32", "b_2": "float32", "c_3": "float32"})  # float32 + int16 -> float32 rows
    for soft in (False, True):
        m.setSoftVoting(soft)
        m.setWeightedTrees(soft)
        for f in (f32, X.astype("float32"), X):
            assert probs_equal(m.testForestProbs(f), orc.forest_probs(trees, scores, m.class_vals, f, soft, soft))
        row = f32.iloc[3]
        assert m.useForest(row) == orc.forest_row(trees, scores, m.class_vals, row, soft, soft)
    m.setSoftVoting(False)
    m.setWeightedTrees(False)
    assert m.testForestProbs(f32) != m.testForestProbs(f32.astype("float64"))  # not vacuous


def test_pickle_and_deepcopy():
    model_checks.check_pickle_and_deepcopy()
