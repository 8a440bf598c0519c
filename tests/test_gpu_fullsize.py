"""GPU (B200), BASELINE.json's full C2 shape (1M rows x 64 mixed features, 4 classes, 8000 + 2000
rows per candidate): size-independent properties, since the CPU oracle cannot finish these sizes.

* conservation: the leaves' class counts of every candidate add up to its bootstrap rows, class by
  class; every leaf is pure unless depth- or rotation-limited; children partition their parent;
* two independent kernels agree: the grower's fused hold-out scorer == the vote kernel run on the
  same tree and the same rows (through the public single-tree validation call);
* idempotence / batching independence: the same draws give bit-identical node tables whether 8 or
  1 candidates share a launch, and a seeded fit gives the same forest for any launch cap;
* the vote of the fitted forest on its own training rows is the same through host-encoded codes
  and through the GPU frame encoder."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    import logging

    import bench
    from random_forest_mc_b200.model import RandomForestMC

    logging.disable(logging.WARNING)
    frame = bench.make_frame(1_000_000, 64, 4, seed=0)
    m = RandomForestMC(n_trees=4, target_col="target", max_discard_trees=2, batch_train_pclass=2000, batch_val_pclass=500)
    m.process_dataset(frame.copy())
    return frame, m


def test_conservation_and_purity(big):
    from random_forest_mc_b200 import engine as eng

    frame, m = big
    np.random.seed(5)
    random.seed(5)
    plan = m._draw_plan(2, 2)
    batch, first = m._grow_plan(plan)
    correct, n_nodes = batch.results()
    trees = batch.fetch(list(range(batch.n_cand)))
    labels = m.encoded.labels
    for ci, t in enumerate(trees):
        slot = ci // 2
        nodes, counts = t.nodes, t.counts
        leaf = nodes["feat"] < 0
        want = np.bincount(labels[m._slot_rows(plan[slot])[0]], minlength=4)
        assert np.array_equal(counts[leaf].sum(axis=0), want)          # rows conserved per class
        assert np.array_equal(counts[0], want)                         # root sees the whole bootstrap
        split = np.flatnonzero(~leaf)
        ch = nodes["child"][split]
        assert np.array_equal(counts[split], counts[ch] + counts[ch + 1])  # children partition parent
        assert (counts[ch].sum(axis=1) > 0).all() and (counts[ch + 1].sum(axis=1) > 0).all()
        pure = (counts[leaf] > 0).sum(axis=1) == 1
        assert pure.mean() > 0.999                                     # grown to purity
        assert 0 <= correct[ci] <= batch.n_val
    # batching independence: candidate 3 alone == candidate 3 inside the batch
    rows1 = m._slot_rows(plan[1])
    solo = eng.GrowBatch(m._dev_dataset, rows1[0][None, :], rows1[1][None, :], [0], [plan[1][2][1]], m.max_depth, 1).run()
    c1, _ = solo.results()
    t1 = solo.fetch([0])[0]
    assert c1[0] == correct[3]
    assert np.array_equal(t1.nodes, trees[3].nodes) and np.array_equal(t1.counts, trees[3].counts)
    solo.close()
    batch.close()


def test_fused_scorer_equals_vote_kernel(big):
    frame, m = big
    np.random.seed(6)
    random.seed(6)
    plan = m._draw_plan(1, 2)
    batch, _ = m._grow_plan(plan)
    correct, _ = batch.results()
    for k in range(2):
        tree = m._wrap(batch.fetch([k])[0])
        score = m.validationTree(tree, m._slot_rows(plan[0])[1])                    # vote kernel on raw values
        assert float(score) == correct[k] / batch.n_val                # fused scorer in the grower
    batch.close()


def test_fit_is_independent_of_launch_cap_and_frame_encoder_agrees(big):
    from random_forest_mc_b200.flat import encode_frame

    frame, m = big
    forests = []
    for cap in (8, 3):
        m.reset_forest()
        m.max_candidates_per_launch = cap
        np.random.seed(7)
        random.seed(7)
        m.fit(disable_progress_bar=True)
        forests.append([(t.soa.nodes.copy(), t.soa.counts.copy(), float(t.survived_score)) for t in m.data])
    for (n1, c1, s1), (n2, c2, s2) in zip(*forests):
        assert s1 == s2 and np.array_equal(n1, n2) and np.array_equal(c1, c2)
    part = frame.drop(columns=["target"]).iloc[:50_000]
    for soft, weighted in ((False, False), (True, True)):
        m.setSoftVoting(soft)
        m.setWeightedTrees(weighted)
        dev = m._device_forest()
        codes, absent = encode_frame(m._forest_tables, part)
        p_codes, l_codes = dev.predict(codes, absent, soft, weighted)
        p_frame, l_frame = dev.predict_frame(part, soft, weighted)
        assert np.array_equal(p_codes, p_frame) and np.array_equal(l_codes, l_frame)
    m.setSoftVoting(False)
    m.setWeightedTrees(False)
    acc = np.mean(np.array(m.testForest(part)) == frame["target"].to_numpy()[:50_000])
    assert acc > 0.27  # 4 trees of median splits on a noisy 4-class target: above chance (0.25)


def test_benched_shape_candidates_equal_the_code_model(big):
    """Two candidates of the bench's own shape (C2b: 8,000 bootstrap rows + 2,000 hold-out rows out of the
    1M x 64 frame, up to 64 features) array-for-array against oracle/code_model.py, scores included."""
    from oracle import code_model as cm

    frame, m = big
    np.random.seed(11)
    random.seed(11)
    plan = m._draw_plan(1, 2)
    batch, first = m._grow_plan(plan)
    correct, n_nodes = batch.results()
    trees = batch.fetch([0, 1])
    idx_T, idx_V = m._slot_rows(plan[0])
    assert len(idx_T) == 8000 and len(idx_V) == 2000
    for k in range(2):
        nodes, counts = cm.grow_codes(m.encoded, idx_T, plan[0][2][k], m.max_depth, m.min_samples_split)
        assert np.array_equal(trees[k].nodes, nodes) and np.array_equal(trees[k].counts, counts)
        assert correct[k] == cm.score_codes(m.encoded, nodes, counts, idx_V)
    batch.close()
