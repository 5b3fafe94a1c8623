"""GPU numerics tests: every hand-written kernel vs a plain PyTorch fp32 reference of the same op."""
import pytest

pytestmark = pytest.mark.gpu


def _checks():
    import kernel_checks
    return kernel_checks.ALL


def pytest_generate_tests(metafunc):
    if "check_name" in metafunc.fixturenames:
        import kernel_checks
        metafunc.parametrize("check_name", [n for n, _ in kernel_checks.ALL])


def test_kernel(check_name):
    import torch
    fn = dict(_checks())[check_name]
    fn()
    torch.cuda.synchronize()
