"""GPU: hypothesis-generated states through the C ABI of libcolosseum_b200.so vs the oracle (tests/cases_property.py)."""
import pytest

import backends
import cases_property as cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    return backends.Cuda()


def test_tron_property(be):
    cases.tron_property(be, examples=150)


def test_ttt_property(be):
    cases.ttt_property(be, examples=150)


def test_blokus_property(be):
    cases.blokus_property(be, examples=40)
