"""The reference's own small known-answer tests (tests/reference_kats.py) through the oracle (CPU) and through the CUDA
engine (GPU): the only per-operator vectors the reference holds besides the TPC-H snapshots."""
import pytest

from tests import reference_kats
from tests.util import gpu_op, oracle_op


def test_reference_kats_oracle():
    reference_kats.check(reference_kats.run(oracle_op))


@pytest.mark.gpu
def test_reference_kats_gpu():
    reference_kats.check(reference_kats.run(gpu_op))
