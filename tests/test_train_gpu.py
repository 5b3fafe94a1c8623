"""GPU parity of the differentiable U-Net step (forward with soft gates + explicit backward to the gates)."""
import pytest

pytestmark = pytest.mark.gpu


def test_train_grads_tiny_with_block_taps():
    import train_checks as T
    T.assert_train(T.check_train_grads())


def test_train_grads_tiny_prediction_only():
    import train_checks as T
    T.assert_train(T.check_train_grads(B=3, use_taps=False))


def test_pruning_step_losses_and_grads_vs_oracle():
    """BASELINE config 3 at test size: six loss scalars within rel 2e-2, d loss/d codebook and d loss/d hypernet
    vs oracle autograd."""
    import train_checks as T
    T.assert_step(*T.check_pruning_step())


def test_finetune_backward_all_parameter_gradients_match_oracle_autograd():
    import train_checks as T
    out = T.check_finetune_grads()
    T.assert_finetune(out)


def test_finetune_step_losses_and_gradients_match_oracle():
    import train_checks as T
    lg, lr, gm, moved = T.check_finetune_step()
    floor = {"distillation_loss": 3e-4, "block_loss": 3e-3}
    for k in lr:
        assert abs(lg[k] - lr[k]) <= 2e-2 * max(abs(lr[k]), 1e-3) + floor.get(k, 0.0), f"{k}: got {lg[k]:.6g} ref {lr[k]:.6g}"
    T.assert_finetune(gm)
    assert not moved, f"parameters of depth-dropped blocks moved: {moved[:5]}"


def test_finetune_of_pruned_expert_matches_physically_pruned_oracle():
    import train_checks as T
    import unet_checks as U
    out, outside, fwd = T.check_finetune_pruned_semantics()
    assert fwd[0] <= 2 * U.MAX_ABS_TOL and fwd[1] >= U.COS_TOL - 5e-4, fwd
    T.assert_finetune(out)
    assert outside and max(outside.values()) == 0.0, {k: v for k, v in outside.items() if v != 0.0}


def test_stale_tape_backward_is_refused():
    """two grad-enabled forwards before one backward must raise, not run on the newer tape."""
    import torch
    import unet_checks as U
    from diffusion_pruning_b200.synthetic import split_arch
    model, _ = U.build_pair(True)
    st = model.get_structure()
    dim = sum(w for ws in st["width"] for w in ws) + 14
    sample, t, ctx = U.inputs(2, 16, model.config["cross_attention_dim"])
    arch = (torch.rand(2, dim) * 0.9 + 0.05).cuda().requires_grad_(True)
    model.set_structure(split_arch(arch * 1.0, st))
    y1 = model(sample.cuda(), t.cuda(), ctx.cuda()).sample
    model.set_structure(split_arch(arch * 1.0, st))
    y2 = model(sample.cuda(), t.cuda(), ctx.cuda()).sample
    y2.float().sum().backward()          # newest tape: fine
    with pytest.raises(RuntimeError, match="tape"):
        y1.float().sum().backward()
