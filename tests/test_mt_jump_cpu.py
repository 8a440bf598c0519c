"""CPU: the MT19937 jump polynomials (csrc/mt_jump.cpp) -- t^(Q-1) mod phi applied to a stream's
first 19936 + 624 words gives the window x[Q .. Q+623] that plain sequential generation (numpy's own
generator) reaches after Q words."""
import numpy as np
import pytest

from random_forest_mc_b200 import engine


def numpy_window(key0, Q):
    """x[Q .. Q+623] of the untempered sequence whose block 0 is key0, through numpy's generator."""
    def block(b):
        if b == 0:
            return key0.copy()
        r = np.random.RandomState()
        r.set_state(("MT19937", key0.copy(), 624, 0, 0.0))
        r.randint(0, 2**32, size=624 * b, dtype=np.uint32)  # full-range uint32: one word per value
        return r.get_state()[1].copy()

    b = Q // 624
    both = np.concatenate([block(b), block(b + 1)])
    return both[Q - 624 * b : Q - 624 * b + 624]


@pytest.mark.parametrize("Q", [0, 1, 5, 623, 624, 625, 19936, 19937, 20561, 624 * 1680, 2 * 624 * 1680, 31 * 624 * 1680 + 77])
def test_jump_window_equals_sequential_generation(Q):
    for seed in (1, 20240229):
        key0 = np.random.RandomState(seed).get_state()[1].copy()
        assert np.array_equal(engine.mt_jump_window_host(key0, Q), numpy_window(key0, Q))


def test_far_jumps_compose():
    """A jump of 3.5e9 words (the 10M-row config consumes that per plan) equals two shorter ones."""
    key0 = np.random.RandomState(4).get_state()[1].copy()
    a, b = 1_234_567_890, 2_265_432_111
    direct = engine.mt_jump_window_host(key0, a + b)
    two = engine.mt_jump_window_host(engine.mt_jump_window_host(key0, a), b)
    assert np.array_equal(direct, two)
