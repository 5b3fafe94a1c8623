"""GPU parity of the routed sampling loop (BASELINE configs[3] at test size)."""
import pytest

pytestmark = pytest.mark.gpu


def test_cfg_ddim_kernel_and_timesteps():
    import sampling_checks as SC
    SC.check_cfg_ddim_kernel()


def test_sampling_loop_vs_oracle():
    import sampling_checks as SC
    max_abs, cos = SC.check_sampling_loop()
    # 4 guided steps (guidance 7.5 amplifies the bf16 error of each step ~7x): looser than the single-step bound
    assert max_abs <= 6e-2 and cos >= 0.999, (max_abs, cos)
