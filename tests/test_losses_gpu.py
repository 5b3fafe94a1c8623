import pytest

pytestmark = pytest.mark.gpu


def test_add_noise_velocity_matches_oracle():
    import loss_checks as LC
    for vp, (e_noisy, e_target) in LC.check_add_noise().items():
        assert e_noisy <= 2e-6 and e_target <= 2e-6, (vp, e_noisy, e_target)  # fp32, one fma vs mul+add


def test_block_mse_matches_fp32_autograd():
    import loss_checks as LC
    for name, (e_loss, e_grad) in LC.check_block_mse().items():
        assert e_loss <= 1e-5, (name, e_loss)          # fp32 differences of bf16 inputs, fp64 accumulation
        assert e_grad <= 8e-3, (name, e_grad)          # gradient is rounded to bf16 (2^-8 relative)


def test_prediction_losses_match_oracle():
    import loss_checks as LC
    for gamma, (e_l, e_d, e_g, e_w) in LC.check_pred_losses().items():
        assert e_l <= 1e-5 and e_d <= 1e-5 and e_g <= 1e-5 and e_w <= 1e-6, (gamma, e_l, e_d, e_g, e_w)


def test_macs_kernel_matches_closed_form_and_oracle():
    import loss_checks as LC
    for other, errs in LC.check_macs_kernel().items():
        # fp64 accumulation in the kernel vs fp32 torch sums of ~1e9..1e11 MAC terms
        assert all(e <= 2e-6 for e in errs), (other, errs)


def test_hypernet_product_matches_reference_golden_and_autograd():
    import loss_checks as LC
    LC.check_hypernet_product_vs_reference_golden()


def test_contrastive_kernels_match_reference_golden_and_autograd():
    import loss_checks as LC
    LC.check_contrastive_kernel_vs_reference_golden()
