"""CPU unit tests: the Tic Tac Toe kernel SOURCE (csrc/ttt.cuh) on the SIMT emulator vs oracle / golden vectors."""
import pytest

import backends
import cases_ttt as cases


@pytest.fixture(scope="module")
def be():
    return backends.HostSim()


def test_line_tables(be):
    cases.case_line_tables(be)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_golden_steps(be, n):
    cases.case_golden_steps(be, n)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_rollout_vs_oracle(be, n):
    cases.case_rollout_vs_oracle(be, n, B=300, K=45)


def test_masked_reset_and_errors(be):
    cases.case_masked_reset_and_errors(be)


@pytest.mark.parametrize("n", [2, 3, 4])
def test_arbitrary_states_reference_recorded(be, n):
    cases.case_arbitrary_states(be, n)
