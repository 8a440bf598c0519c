"""GPU: the device walk of numpy's legacy stream (rfmc_draw_plan_device, csrc/rng_walk.cu) draws the
same rows and feature counts and leaves the stream in the same state as numpy itself
(split_train_val / sampleFeats, model.py:198-219) and as the host restatement (host_rng.cpp)."""
import numpy as np
import pytest

from random_forest_mc_b200 import engine

pytestmark = pytest.mark.gpu


def make_dataset(sizes, seed=0):
    """A dataset whose classes have the given sizes (rows interleaved at random)."""
    g = np.random.default_rng(seed)
    labels = np.concatenate([np.full(s, str(c)) for c, s in enumerate(sizes)])
    g.shuffle(labels)
    n = len(labels)
    cols = {"x": (np.arange(n) % 7).astype(np.float64), "target": labels.astype(object)}
    class_vals = list(dict.fromkeys(labels.tolist()))
    enc, ds = engine.encode_dataset_device(cols, ["x"], ["x"], "target", class_vals)
    return enc, ds


def numpy_plan(class_rows, n_slots, per_class, n_train, n_cands, lo, hi):
    T, V, K, states = [], [], [], []
    for _ in range(n_slots):
        states.append(np.random.get_state())
        tr, va = [], []
        for rows in class_rows:
            sel = np.random.choice(rows, per_class, replace=False)
            tr.extend(sel[:n_train])
            va.extend(sel[n_train:])
        T.append(tr), V.append(va)
        K.append([np.random.randint(lo, hi) for _ in range(n_cands)])
    return np.array(T, dtype=np.int64), np.array(V, dtype=np.int64), np.array(K), states, np.random.get_state()


CASES = [
    # sizes, n_slots, per_class, n_train, n_cands, lo, hi
    ([1, 2, 3], 3, 1, 1, 2, 2, 9),
    ([2, 2], 4, 2, 1, 1, 0, 1),  # randint over one value: no draw
    ([5, 9, 17, 33], 6, 4, 2, 3, 2, 5),
    ([64, 65, 127, 128, 129], 5, 20, 10, 4, 2, 65),
    ([700, 1501, 333, 4096], 5, 20, 10, 3, 2, 65),
    ([1023, 1024, 1025, 2047, 2048, 2049], 3, 300, 250, 2, 3, 9),
    ([4095, 4097, 8191, 8193], 3, 7, 10, 2, 2, 3),
    ([16383, 16385, 32767, 32769], 3, 50, 40, 4, 2, 130),
    ([65535, 65537, 70001], 2, 600, 500, 4, 2, 65),
    ([131071, 131073, 200000], 2, 25, 10, 2, 5, 6),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_device_plan_matches_numpy(case):
    sizes, n_slots, per_class, n_train, n_cands, lo, hi = CASES[case]
    enc, ds = make_dataset(sizes, seed=case)
    try:
        for seed, skip in ((11, 3), (5, 0), (99, 311)):
            np.random.seed(seed)
            np.random.random_sample(skip)  # an arbitrary position inside the key block
            st = np.random.get_state()
            T, V, K, states, after = numpy_plan(enc.class_rows, n_slots, per_class, n_train, n_cands, lo, hi)
            plan = engine.DrawPlan(ds, enc.class_rows, st[1], int(st[2]), n_slots, per_class, n_train, n_cands, lo, hi)
            try:
                counts, key_end, pos_end = plan.finish()
                gT, gV = plan.indices()
                assert np.array_equal(gT, T.reshape(gT.shape)), "training rows differ"
                assert np.array_equal(gV, V.reshape(gV.shape)), "hold-out rows differ"
                assert np.array_equal(counts, K.reshape(counts.shape)), "randint values differ"
                assert np.array_equal(key_end, after[1]) and pos_end == after[2], "end state differs"
                for s in range(n_slots):
                    k, p = plan.state_at(s)
                    assert np.array_equal(k, states[s][1]) and p == states[s][2], f"state before slot {s} differs"
            finally:
                plan.close()
    finally:
        ds.close()


def test_device_plan_at_full_size_matches_host_walk():
    """1M rows, 4 classes (the bench shape): rows, counts, snapshots and end state equal the host
    restatement's, which tests/test_host_rng.py pins to numpy; also through several fill chunks."""
    sizes = [250_000, 250_001, 249_999, 250_000]
    enc, ds = make_dataset(sizes, seed=3)
    try:
        for per_class, n_train, n_slots in ((20, 10, 12), (2500, 2000, 4)):
            np.random.seed(21)
            np.random.random_sample(5)
            st = np.random.get_state()
            key = st[1].copy()
            T, V, K, ksnap, psnap, pos = engine.legacy_draw_plan(key, int(st[2]), n_slots, enc.class_rows, per_class, n_train, 4, 2, 65, 2)
            plan = engine.DrawPlan(ds, enc.class_rows, st[1], int(st[2]), n_slots, per_class, n_train, 4, 2, 65)
            try:
                counts, key_end, pos_end = plan.finish()
                gT, gV = plan.indices()
                assert np.array_equal(gT, T) and np.array_equal(gV, V)
                assert np.array_equal(counts, K)
                assert np.array_equal(key_end, key) and pos_end == pos
                for s in (0, 1, n_slots - 1):
                    k, p = plan.state_at(s)
                    assert np.array_equal(k, ksnap[s]) and p == psnap[s]
                info = plan.info()
                assert info["words_consumed"] > 0 and info["words_generated"] >= info["words_consumed"]
            finally:
                plan.close()
    finally:
        ds.close()


def test_every_draw_of_a_long_plan_matches_the_host_walk():
    """256 slots x 4 classes of 250k rows, 2,500 rows drawn per class: 1,024 resolves with ~12k chain
    hops each (two steps of one tile hitting the same live head, heads born and hit inside one tile,
    table rebuilds) -- every row id equal to the host restatement's."""
    sizes = [250_000, 250_001, 249_999, 250_000]
    enc, ds = make_dataset(sizes, seed=0)
    try:
        np.random.seed(1)
        st = np.random.get_state()
        key = st[1].copy()
        T, V, K, ksnap, psnap, pos = engine.legacy_draw_plan(key, int(st[2]), 256, enc.class_rows, 2500, 2000, 4, 2, 65, 2)
        plan = engine.DrawPlan(ds, enc.class_rows, st[1], int(st[2]), 256, 2500, 2000, 4, 2, 65)
        try:
            counts, key_end, pos_end = plan.finish()
            gT, gV = plan.indices()
            assert np.array_equal(gT, T) and np.array_equal(gV, V)
            assert np.array_equal(counts, K) and np.array_equal(key_end, key) and pos_end == pos
        finally:
            plan.close()
    finally:
        ds.close()


def test_plan_feeds_the_grower_without_a_host_copy():
    """rfmc_batch_create_from_plan == rfmc_batch_create on the same rows."""
    g = np.random.default_rng(5)
    n = 30_000
    cols = {"a": np.round(g.normal(size=n), 2), "b": g.integers(0, 50, size=n), "target": g.choice(["u", "v", "w"], size=n).astype(object)}
    class_vals = list(dict.fromkeys(cols["target"].tolist()))
    enc, ds = engine.encode_dataset_device(cols, ["a", "b"], ["a", "b"], "target", class_vals)
    try:
        np.random.seed(2)
        st = np.random.get_state()
        plan = engine.DrawPlan(ds, enc.class_rows, st[1], int(st[2]), 3, 60, 40, 2, 1, 3)
        plan.finish()
        T, V = plan.indices()
        feats = [[0, 1], [1], [1, 0], [0], [0, 1], [1, 0]]
        slot_of = [0, 0, 1, 1, 2, 2]
        a = engine.GrowBatch(ds, None, None, slot_of, feats, 1000, 1, plan=plan).run()
        b = engine.GrowBatch(ds, T, V, slot_of, feats, 1000, 1).run()
        ca, na = a.results()
        cb, nb = b.results()
        assert np.array_equal(ca, cb) and np.array_equal(na, nb)
        for x, y in zip(a.fetch(range(6)), b.fetch(range(6))):
            assert np.array_equal(x.nodes, y.nodes) and np.array_equal(x.counts, y.counts)
        a.close(), b.close(), plan.close()
    finally:
        ds.close()


def test_fit_with_device_draws_equals_fit_with_host_draws():
    """fit() with numpy's stream walked on the GPU (device_draws) grows the forest of the host walk and
    leaves both RNG streams in the same state; a plan that stops early exercises the rewind through the
    device plan's stream states."""
    import random

    import pandas as pd

    from random_forest_mc_b200.model import RandomForestMC
    from tests.helpers import tree_diff

    g = np.random.default_rng(8)
    n = 6000
    frame = pd.DataFrame({"a": np.round(g.normal(size=n), 2), "b": g.integers(0, 30, size=n), "c": g.choice(list("xyz"), size=n)})
    frame["target"] = ((frame.a + frame.b / 15 + (frame.c == "x") + g.normal(size=n) * 0.7) > 1.3).astype(int).astype(str)
    out = []
    for device in (False, True):
        for th in (1.0, 0.55):  # 0.55: most slots stop at their first candidate -> rewinds
            m = RandomForestMC(n_trees=12, target_col="target", max_discard_trees=3, batch_train_pclass=40, batch_val_pclass=30, th_start=th)
            m.device_draws = device
            m.process_dataset(frame.copy())
            np.random.seed(3), random.seed(3)
            m.fit()
            out.append((m, np.random.get_state(), random.getstate()))
    for a, b in ((out[0], out[2]), (out[1], out[3])):
        assert len(a[0].data) == len(b[0].data) == 12
        for ta, tb in zip(a[0].data, b[0].data):
            assert tree_diff(ta.data, tb.data) is None
        assert a[0].survived_scores == b[0].survived_scores
        assert np.array_equal(a[1][1], b[1][1]) and a[1][2] == b[1][2] and a[2] == b[2]
    assert out[3][0].fit_stats["rewinds"] > 0
