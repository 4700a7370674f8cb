"""The reference's hot-path unit tests against the CPU oracle, the unmodified reference
(when oracle/_ref was built) and — marked gpu — the CUDA library through the C ABI."""
import pytest

from tests.backends import CudaBackend
from tests.reference_cases import CASES


@pytest.mark.parametrize("case", CASES, ids=[c.__name__ for c in CASES])
def test_oracle(oracle, case):
    case(oracle)


@pytest.mark.parametrize("case", CASES, ids=[c.__name__ for c in CASES])
def test_unmodified_reference(reference, case):
    case(reference)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c.__name__ for c in CASES])
def test_cuda(gpu, case):
    case(CudaBackend(gpu))
