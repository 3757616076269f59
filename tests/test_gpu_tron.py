"""GPU parity tests (B200): Tron through the C ABI of libcolosseum_b200.so vs the oracle / golden vectors."""
import numpy as np
import pytest

import backends
import cases_tron as cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    return backends.Cuda()


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


def test_rollout_vs_oracle_small(be):
    cases.case_rollout_vs_oracle(be, N=19, P=4, B=130, K=30)
    cases.case_rollout_vs_oracle(be, N=7, P=3, B=40, K=30, seed=9)
    cases.case_rollout_vs_oracle(be, N=8, P=2, B=1000, K=40, seed=2)


def test_rollout_vs_oracle_full_size(be):
    # BASELINE.json configs[1]: 65,536 environments, 19x19, 4 players; end states + episode statistics bit-exact
    cases.case_rollout_vs_oracle(be, N=19, P=4, B=65536, K=64, seed=0, env0=0)


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
    # shapes beyond N <= 19, P <= 4 (csrc/tron_wide.cuh): policy + step + fused rollout + statistics vs the oracle
    cases.case_rollout_vs_oracle(be, N=21, P=4, B=300, K=60, seed=4)
    cases.case_rollout_vs_oracle(be, N=11, P=6, B=500, K=40, seed=5)
    cases.case_rollout_vs_oracle(be, N=25, P=8, B=2000, K=50, seed=6, env0=77)
    cases.case_rollout_vs_oracle(be, N=64, P=8, B=70, K=90, seed=7)            # the largest shape: 128 words per plane
    cases.case_rollout_vs_oracle(be, N=20, P=2, B=129, K=120, seed=8)          # the smallest wide shape


def test_wide_adversarial(be):
    cases.case_wide_adversarial(be)


def test_wide_in_place_and_masked_reset(be):
    cases.case_in_place_and_masked_reset(be, N=21, P=4)
    cases.case_in_place_and_masked_reset(be, N=11, P=6, B=130)
