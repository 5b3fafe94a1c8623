"""GPU parity of the router kernels: golden vectors from the reference's own code, then BASELINE config 5
(4096 prompts x 8 codes) against the CPU oracle -- assignments must be bit-exact."""
import pytest

pytestmark = pytest.mark.gpu


def test_router_matches_reference_goldens():
    import router_checks as RC
    RC.check_router_vs_golden()


def test_router_config5_4096x8_bit_exact():
    import router_checks as RC
    assert RC.check_router_vs_oracle(batch=4096)


def test_router_edge_batches():
    import router_checks as RC
    for b in (1, 7, 130):
        assert RC.check_router_vs_oracle(batch=b, seed=10 + b)


def test_router_backward():
    import router_checks as RC
    RC.check_router_backward()
