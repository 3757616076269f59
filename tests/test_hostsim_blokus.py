"""CPU unit tests: the Blokus kernel SOURCE (csrc/blokus.cuh, warp-level code) on the SIMT emulator vs oracle / golden."""
import pytest

import backends
import cases_blokus as cases


@pytest.fixture(scope="module")
def be():
    return backends.HostSim()


def test_reset_and_capacity(be):
    cases.case_reset_and_capacity(be)


def test_golden_games(be):
    cases.case_golden_games(be)


def test_illegal_actions(be):
    cases.case_illegal_actions(be)


def test_rollout_vs_oracle(be):
    cases.case_rollout_vs_oracle(be, B=4, K=72)


def test_random_boards(be):
    cases.case_random_boards(be, n=20)


def test_many_anchors(be):
    cases.case_many_anchors(be)


def test_is_valid(be):
    cases.case_is_valid(be, stride=6)


def test_wide_golden(be):
    cases.case_wide_golden(be, stride=9)


def test_arbitrary_positions_reference_recorded(be):
    cases.case_arbitrary_positions(be)
