"""GPU numerics of the backward kernels (K5) vs torch autograd of the plain fp32 ops."""
import pytest

pytestmark = pytest.mark.gpu


def pytest_generate_tests(metafunc):
    if "bwd_name" in metafunc.fixturenames:
        import backward_checks
        metafunc.parametrize("bwd_name", [n for n, _ in backward_checks.ALL])


def test_backward_kernel(bwd_name):
    import torch
    import backward_checks
    dict(backward_checks.ALL)[bwd_name]()
    torch.cuda.synchronize()
