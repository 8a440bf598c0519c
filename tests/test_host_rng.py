"""CPU: the host restatement of numpy's legacy choice(replace=False) draws the same indices and
leaves the MT19937 stream in the same state as numpy itself."""
import numpy as np
import pytest

from random_forest_mc_b200 import engine


@pytest.mark.parametrize("n,size", [(1, 1), (2, 2), (3, 1), (17, 5), (256, 20), (257, 257), (1000, 20), (65536, 40), (250001, 2500)])
def test_choice_matches_numpy(n, size):
    a = np.arange(n) * 3 + 7
    for seed in (0, 1, 12345):
        np.random.seed(seed)
        np.random.random_sample(seed % 7)  # start from an arbitrary position inside the key block
        st = np.random.get_state()
        want = np.random.choice(a, size, replace=False)
        after = np.random.get_state()
        key = st[1].copy()
        idx, pos = engine.legacy_choice_indices(key, int(st[2]), n, size)
        assert np.array_equal(a[idx], want)
        assert np.array_equal(key, after[1]) and pos == after[2]


def test_stream_continues_exactly():
    np.random.seed(5)
    a = np.arange(5000)
    want = [np.random.choice(a, 30, replace=False) for _ in range(3)] + [np.random.randint(2, 60)]
    np.random.seed(5)
    got = []
    for _ in range(3):
        st = np.random.get_state()
        key = st[1].copy()
        idx, pos = engine.legacy_choice_indices(key, int(st[2]), len(a), 30)
        np.random.set_state((st[0], key, pos, st[3], st[4]))
        got.append(a[idx])
    got.append(np.random.randint(2, 60))
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


@pytest.mark.parametrize("threads", [0, 3, 0x103])
@pytest.mark.parametrize("per_class,n_train,lo,hi", [(20, 10, 2, 65), (7, 10, 2, 3), (300, 250, 3, 9), (5, 5, 4, 5)])
def test_draw_plan_matches_numpy_call_sequence(threads, per_class, n_train, lo, hi):
    """choice x classes then randint x candidates per slot, exactly numpy's values and end state."""
    g = np.random.default_rng(3)
    sizes = [700, 1501, 333, 4096]
    class_rows = [np.sort(g.choice(100000, s, replace=False)).astype(np.int64) for s in sizes]
    n_slots, n_cands = 5, 3
    np.random.seed(11)
    np.random.random_sample(3)
    st = np.random.get_state()
    want_T, want_V, want_k, want_state = [], [], [], []
    for s in range(n_slots):
        want_state.append(np.random.get_state())
        tr, va = [], []
        for rows in class_rows:
            sel = np.random.choice(rows, per_class, replace=False)
            tr.extend(sel[:n_train])
            va.extend(sel[n_train:])
        want_T.append(tr), want_V.append(va)
        want_k.append([np.random.randint(lo, hi) for _ in range(n_cands)])
    after = np.random.get_state()
    key = st[1].copy()
    T, V, K, ksnap, psnap, pos = engine.legacy_draw_plan(key, int(st[2]), n_slots, class_rows, per_class, n_train, n_cands, lo, hi, threads)
    assert np.array_equal(T, np.array(want_T, dtype=np.int64).reshape(T.shape))
    assert np.array_equal(V, np.array(want_V, dtype=np.int64).reshape(V.shape))
    assert np.array_equal(K, np.array(want_k))
    assert np.array_equal(key, after[1]) and pos == after[2]
    for s in range(n_slots):
        assert np.array_equal(ksnap[s], want_state[s][1]) and psnap[s] == want_state[s][2]


@pytest.mark.parametrize("n_pop", [1, 2, 6, 21, 22, 64, 85, 86, 128, 1000])
def test_pyrandom_sample_matches_cpython(n_pop):
    """Indices and end state of random.sample, for the module-level stream and a Random instance,
    across the pool / set branches (setsize = 21 + 4**ceil(log(3k, 4)))."""
    import random

    g = np.random.default_rng(n_pop)
    ks = np.minimum(g.integers(0, max(2, n_pop + 1), size=40), n_pop).astype(np.int32)
    ks[:6] = [0, min(1, n_pop), min(5, n_pop), min(6, n_pop), n_pop, min(7, n_pop)]
    pop = list(range(n_pop))
    for make in (lambda: random, lambda: random.Random(77)):
        rng = make()
        rng.seed(1234 + n_pop)
        rng.random()
        want = [rng.sample(pop, int(k)) for k in ks]
        after = rng.getstate()
        rng.seed(1234 + n_pop)
        rng.random()
        got = engine.pyrandom_sample_plan(rng, n_pop, ks)
        assert rng.getstate() == after
        at = 0
        for k, w in zip(ks, want):
            assert got[at : at + k].tolist() == w
            at += int(k)
