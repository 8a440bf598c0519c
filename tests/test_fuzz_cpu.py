"""CPU: the randomized parity cases of tests/test_gpu_fuzz.py on the oracle-backed fake engine
(exercises the encoder, the planner/rewind logic, flattening and frame encoding on odd inputs)."""
import pytest

from tests import fake_engine
from tests.test_gpu_fuzz import run_case


@pytest.fixture(autouse=True)
def _fake(monkeypatch):
    fake_engine.install(monkeypatch)


@pytest.mark.parametrize("seed", list(range(100, 112)) + [1001, 1304])  # the last two: empty hold-out set -> the reference's IndexError
def test_random_config_matches_oracle_on_fake_engine(seed):
    run_case(seed)


@pytest.mark.parametrize("seed", range(12))
def test_wide_random_config_matches_oracle_on_fake_engine(seed):
    run_case(seed, wide=True)
