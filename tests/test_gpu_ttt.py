"""GPU parity tests (B200): Tic Tac Toe through the C ABI vs the oracle / golden vectors."""
import pytest

import backends
import cases_ttt as cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    return backends.Cuda()


def test_line_tables(be):
    cases.case_line_tables(be)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_golden_steps(be, n):
    cases.case_golden_steps(be, n)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_rollout_vs_oracle(be, n):
    cases.case_rollout_vs_oracle(be, n, B=3000, K=60)


def test_rollout_vs_oracle_full_size(be):
    # BASELINE.json configs[3]: 4-player 3x3x3, 1,048,576 environments
    cases.case_rollout_vs_oracle(be, 4, B=1 << 20, K=48, seed=0, env0=0)


def test_masked_reset_and_errors(be):
    cases.case_masked_reset_and_errors(be)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_arbitrary_states_reference_recorded(be, n):
    cases.case_arbitrary_states(be, n)
