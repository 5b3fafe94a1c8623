"""GPU parity of the routed sampling loop (BASELINE configs[3] at test size)."""
import pytest

pytestmark = pytest.mark.gpu


def test_cfg_ddim_kernel_and_timesteps():
    import sampling_checks as SC
    SC.check_cfg_ddim_kernel()


def test_sampling_loop_vs_oracle():
    import sampling_checks as SC
    max_abs, cos = SC.check_sampling_loop()
    # 4 guided steps: pred = u + 7.5 (c - u) amplifies the bf16 error of each U-Net step (single-step bound:
    # cosine >= 0.9998, i.e. relative error 2e-2) by up to sqrt(6.5^2 + 7.5^2) ~ 10x before the DDIM update mixes
    # it back into x, so the loop bound is looser than the single-step one: cosine >= 0.998 (relative error
    # 6e-2), max-abs <= 6e-2 of the output scale. Measured on B200: cosine 0.99896, max-abs 4.8e-2.
    assert max_abs <= 6e-2 and cos >= 0.998, (max_abs, cos)
