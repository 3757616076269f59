"""GPU parity tests (B200): Blokus through the C ABI vs the oracle / golden vectors."""
import pytest

import backends
import cases_blokus as cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    return backends.Cuda()


def test_reset_and_capacity(be):
    cases.case_reset_and_capacity(be)


def test_golden_games(be):
    cases.case_golden_games(be)


def test_illegal_actions(be):
    cases.case_illegal_actions(be)


def test_rollout_vs_oracle(be):
    cases.case_rollout_vs_oracle(be, B=64, K=150)


def test_rollout_vs_oracle_wide(be):
    # many games, two+ full episodes each: end states and fused statistics bit-exact vs the oracle
    cases.case_rollout_vs_oracle(be, B=1024, K=150, seed=0, env0=0)


def test_rollout_vs_oracle_full_size(be):
    # BASELINE.json configs[2] at its full size: 16 384 games x 70 steps (one full episode and the restart), end
    # states and the fused statistics bit-exact vs the oracle (~20 s of oracle time on the box's host cores)
    cases.case_rollout_vs_oracle(be, B=16384, K=70, seed=2, env0=0, cap=4096)


def test_random_boards(be):
    cases.case_random_boards(be, n=160)


def test_many_anchors(be):
    cases.case_many_anchors(be)


def test_is_valid(be):
    cases.case_is_valid(be)


def test_wide_golden(be):
    cases.case_wide_golden(be)


def test_arbitrary_positions_reference_recorded(be):
    cases.case_arbitrary_positions(be)
