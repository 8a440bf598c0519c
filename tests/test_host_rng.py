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
