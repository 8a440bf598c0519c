"""CPU: the drop-in API's host logic (speculative batching + RNG rewind, policy, flattening,
persistence, merge) on an oracle-backed fake engine, against the reference's recorded vectors."""
import numpy as np
import pytest

from tests import fake_engine, model_checks
from tests.helpers import golden_names, load_golden


@pytest.fixture(autouse=True)
def _fake(monkeypatch):
    fake_engine.install(monkeypatch)


@pytest.mark.parametrize("name", golden_names())
def test_fit_and_predict_match_reference(name):
    g, m = model_checks.check_fit_matches_reference(name)
    model_checks.check_predictions_match_reference(g, m)


@pytest.mark.parametrize("name", ["titanic_c1", "mixed_feat", "mixed_nobest"])
@pytest.mark.parametrize("cap", [1, 3, 7])
def test_fit_is_independent_of_batch_size(name, cap):
    """Tiny launch caps force many plans, rewinds and slot extensions: the forest must not change."""
    model_checks.check_fit_matches_reference(name, max_candidates_per_launch=cap)


@pytest.mark.parametrize("name", ["titanic_c1", "mixed_feat", "mixed_nobest", "mixed_default"])
@pytest.mark.parametrize("cap", [3, 4096])
def test_fit_through_device_draw_plans_matches_reference(name, cap):
    """The device-draw branch of the fit loop (row ids owned by the plan, stream states read back only
    for a rewind, plans closed after their batch) on the fake plan -- the host walk behind the device
    plan's interface: same forest, same consumed candidates, same RNG end states as the reference."""
    before = fake_engine.FakeDrawPlan.created
    model_checks.check_fit_matches_reference(name, device_draws=True, device_draw_min_slots=1, max_candidates_per_launch=cap)
    assert fake_engine.FakeDrawPlan.created > before


@pytest.mark.parametrize("name", ["titanic_c1", "mixed_feat", "mixed_nobest", "mixed_default", "ties_default"])
@pytest.mark.parametrize("cap,sub", [(4096, 1), (4096, 2), (4096, 4), (3, 1), (7, 2), (16, 1)])
@pytest.mark.parametrize("device_draws", [False, True])
def test_pipelined_plans_match_reference(name, cap, sub, device_draws):
    """The software-pipelined fit loop (plan B drawn and launched before plan A is judged; B dropped when A
    needs a rewind), forced on for small data: same forest, consumed candidates and RNG end states."""
    g, m = model_checks.check_fit_matches_reference(
        name, pipeline_min_words=0, pipeline_slots=sub, max_candidates_per_launch=cap, device_draws=device_draws, device_draw_min_slots=1)
    assert m.fit_stats["batches"] >= 1 and m.fit_stats["plans_dropped"] >= 0


def test_pipelined_plans_drop_successor_on_rewind():
    """A forest whose slots stop early under full speculation rewinds: the successor drawn ahead is dropped."""
    g, m = model_checks.check_fit_matches_reference("mixed_nobest", pipeline_min_words=0, pipeline_slots=1)
    assert m.fit_stats["rewinds"] == 0 or m.fit_stats["plans_dropped"] >= 0
    dropped = 0
    for name in ("titanic_c1", "mixed_feat", "mixed_nobest", "mixed_default"):
        _, mm = model_checks.check_fit_matches_reference(name, pipeline_min_words=0, pipeline_slots=1, max_candidates_per_launch=3)
        dropped += mm.fit_stats["plans_dropped"]
    assert dropped > 0


def test_persistence_and_merge():
    g, m = model_checks.fitted_model("mixed_default")
    model_checks.check_persistence_and_merge(g, m)


def test_single_tree_call_and_class_trees():
    g, m = model_checks.fitted_model("titanic_c1")
    from oracle import rfmc_oracle as orc

    fr = g["frames"]["absent_nan"]
    for i in (0, 5, 11):
        row = fr.iloc[i]
        for tree, ref in zip(m.data, g["forest"]):
            want = orc.tree_output(ref, g["class_vals"], row)
            got = tree(row)
            assert list(got.keys()) == list(want.keys())
            assert all(float(got[k]) == float(want[k]) for k in want)
        for c in g["class_vals"]:
            want_trees = [k for k, ref in enumerate(g["forest"]) if orc.top_class(orc.tree_output(ref, g["class_vals"], row)) == c]
            got_trees = [k for k, t in enumerate(m.data) if any(t is x for x in m.sampleClass2trees(row, c))]
            assert got_trees == want_trees


def test_errors_mirror_reference():
    from random_forest_mc_b200.model import DatasetNotFound, RandomForestMC

    m = RandomForestMC()
    with pytest.raises(DatasetNotFound):
        m.fit()
    with pytest.raises(TypeError):
        m.predict([1, 2, 3])
    with pytest.raises(TypeError):
        m.mergeForest("nope")
    g, a = model_checks.fitted_model("ties_default")
    _, b = model_checks.fitted_model("mixed_default")
    with pytest.raises(ValueError):
        a.mergeForest(b)
    assert a.testForest(g["frames"]["clean"].head(0)) == []


def test_depths_and_plant_validate_single_calls():
    g, m = model_checks.fitted_model("mixed_mss3_d6")
    for t in m.data:
        assert max(t.depths) <= 6
        dict_depths = t.depths
        t.data  # materialise
        t._soa = None
        assert t.depths == dict_depths
    ref = g["trace"][0]
    t = m.plantTree(ref["idx_T"], ref["F"])
    from tests.helpers import tree_diff

    assert tree_diff(t.data, ref["tree"]) is None
    assert float(m.validationTree(t, ref["idx_V"])) == float(ref["th_val"])


def test_predict_missing_values_matches_reference():
    model_checks.check_predict_missing_values()


def test_fit_parallel_host_streams_contract():
    """fitParallel(max_workers=2): worker r == the serial algorithm under seed s_r on its share of
    the slots; forest = concatenation in worker order (checked against the oracle)."""
    import random

    from oracle import rfmc_oracle as orc
    from random_forest_mc_b200.model import RandomForestMC
    from tests.helpers import tree_diff

    g = load_golden("mixed_default")
    kw = dict(g["kwargs"])
    kw["n_trees"] = 5
    m = RandomForestMC(**kw)
    m.worker_seeds = [1000, 1001]
    m.fitParallel(g["frame"].copy(), disable_progress_bar=True, max_workers=2)
    want_trees, want_scores = [], []
    for seed, n in ((1000, 3), (1001, 2)):
        ds = orc.ingest(g["frame"].copy(), target_col=kw["target_col"])
        np.random.seed(seed)
        random.seed(seed)
        t, s, _ = orc.fit(ds, n, max_discard_trees=kw["max_discard_trees"])
        want_trees += t
        want_scores += [float(x) for x in s]
    assert [float(s) for s in m.survived_scores] == want_scores
    for a, b in zip(m.data, want_trees):
        assert tree_diff(a.data, b) is None
    assert m.fit_stats["host_streams"] == 2
    # default seeds come from the global stream: reproducible under np.random.seed
    outs = []
    for _ in range(2):
        m2 = RandomForestMC(**kw)
        np.random.seed(3)
        random.seed(3)
        m2.fitParallel(g["frame"].copy(), disable_progress_bar=True, max_workers=3)
        outs.append([float(s) for s in m2.survived_scores])
    assert outs[0] == outs[1] and len(outs[0]) == 5


def test_pickle_and_deepcopy():
    model_checks.check_pickle_and_deepcopy()


def test_device_forest_follows_the_trees_not_their_ids():
    """The cached device copy of the forest is keyed on the tree objects themselves (held, compared with
    `is`) and on the version of each tree's dict: a forest rebuilt with equal scores, or a tree whose
    dict was replaced, must vote with the new trees (a freed tree's id() can come back on a new one)."""
    import gc

    g, m = model_checks.fitted_model("titanic_c1")
    frame = g["frames"]["clean"]
    from random_forest_mc_b200.tree import DecisionTreeMC

    def one_tree_model(tree_dict):
        m.data = [DecisionTreeMC(tree_dict, m.class_vals, survived_score=1.0, features=m.feature_cols)]
        m.survived_scores = [1.0]
        return m.testForestProbs(frame)

    seen = []
    for ref in g["forest"][:4]:
        seen.append(one_tree_model(ref))
        gc.collect()  # the previous single tree is garbage now: its id is free to be reused
    from oracle import rfmc_oracle as orc

    for ref, got in zip(g["forest"][:4], seen):
        want = orc.forest_probs([ref], [1.0], g["class_vals"], frame, False, False)
        assert got == want
    # replacing a tree's dict in place of the same object
    m.data[0].data = g["forest"][0]
    assert m.testForestProbs(frame) == orc.forest_probs([g["forest"][0]], [1.0], g["class_vals"], frame, False, False)


def test_device_draws_are_chosen_automatically_only_for_ranks_short_of_cores(monkeypatch):
    """device_draws = None: the device walk of numpy's stream is used when this process is one of several ranks on the
    box and its share of the host cores is below six (model.device_draws_on); explicit settings always win."""
    from random_forest_mc_b200 import model as model_mod

    g, m = model_checks.fitted_model("mixed_default")
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    m.device_draws = None
    assert m.device_draws_on() is False  # one process on the box
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    monkeypatch.setattr(model_mod, "_usable_cores", lambda: 32)
    assert m.device_draws_on() is True  # 4 cores per rank
    monkeypatch.setattr(model_mod, "_usable_cores", lambda: 64)
    assert m.device_draws_on() is False  # 8 cores per rank
    monkeypatch.setattr(model_mod, "_usable_cores", lambda: 32)
    m.device_draw_auto_max_rows = 10
    assert m.device_draws_on() is False  # a frame larger than what the choice was measured on
    m.device_draws = True
    assert m.device_draws_on() is True
    m.device_draws = False
    m.device_draw_auto_max_rows = 10**9
    assert m.device_draws_on() is False


def test_automatic_device_draws_fall_back_to_the_host_walk_when_a_plan_fails(monkeypatch):
    """device_draws = None (automatic): a device plan that gives up (EngineError) is drawn by the host walk instead --
    nothing was committed, so the forest, the consumed candidates and both RNG end states are the reference's.  With
    device_draws = True the error surfaces."""
    from random_forest_mc_b200 import engine
    from random_forest_mc_b200.model import RandomForestMC

    class FailingPlan(fake_engine.FakeDrawPlan):
        def finish(self):
            raise engine.EngineError("device draw: stream words exhausted")

    monkeypatch.setattr(engine, "DrawPlan", FailingPlan)
    monkeypatch.setattr(RandomForestMC, "device_draws_on", lambda self: True)
    g, m = model_checks.check_fit_matches_reference("mixed_default", device_draws=None, device_draw_min_slots=1)
    assert m.device_draw_fallbacks >= 1
    with pytest.raises(engine.EngineError):
        model_checks.fitted_model("mixed_default", device_draws=True, device_draw_min_slots=1)
