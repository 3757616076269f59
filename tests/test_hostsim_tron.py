"""CPU unit tests: the Tron kernel SOURCE (csrc/tron.cuh) executed on the SIMT emulator vs the oracle and the
golden vectors of the real reference.  (The same cases run on the real GPU in test_gpu_tron.py.)"""
import pytest

import backends
import cases_tron as cases


@pytest.fixture(scope="module")
def be():
    return backends.HostSim()


def test_start_positions(be):
    cases.case_start_positions(be)


def test_reset(be):
    cases.case_reset(be)


@pytest.mark.parametrize("path", cases._golden_files(), ids=lambda p: p.split("/")[-1])
def test_golden_steps(be, path):
    cases.case_golden_steps(be, path)


def test_adversarial(be):
    cases.case_adversarial(be)


def test_adversarial_wide_reference_recorded(be):
    cases.case_adversarial(be, "tron_adversarial_wide.npz")


def test_rollout_vs_oracle(be):
    cases.case_rollout_vs_oracle(be, N=19, P=4, B=130, K=30)
    cases.case_rollout_vs_oracle(be, N=7, P=3, B=40, K=30, seed=9)


def test_in_place_and_masked_reset(be):
    cases.case_in_place_and_masked_reset(be)


def test_errors(be):
    cases.case_errors(be)


def test_compact_result(be):
    cases.case_compact_result(be)


def test_packed_actions(be):
    cases.case_packed_actions(be)


def test_start_positions_golden(be):
    cases.case_start_positions_golden(be)


def test_wide_rollout_vs_oracle(be):
    cases.case_rollout_vs_oracle(be, N=21, P=4, B=40, K=40, seed=4)
    cases.case_rollout_vs_oracle(be, N=11, P=6, B=70, K=30, seed=5)
    cases.case_rollout_vs_oracle(be, N=64, P=8, B=8, K=40, seed=7)             # the largest shape: 128 words per plane


def test_wide_adversarial(be):
    cases.case_wide_adversarial(be, n=120)


def test_wide_in_place_and_masked_reset(be):
    cases.case_in_place_and_masked_reset(be, N=11, P=6, B=40)
