"""CPU: hypothesis-generated states, the kernel SOURCE on the SIMT emulator vs the oracle (tests/cases_property.py)."""
import pytest

import backends
import cases_property as cases


@pytest.fixture(scope="module")
def be():
    return backends.HostSim()


def test_tron_property(be):
    cases.tron_property(be, examples=30)


def test_ttt_property(be):
    cases.ttt_property(be, examples=30)


def test_blokus_property(be):
    cases.blokus_property(be, examples=8)
