"""Training-path parity: the differentiable (soft-gate) CUDA U-Net and its explicit backward vs autograd through
the fp32 CPU oracle on identical seeded weights, inputs and gates. Gradients reach only the gate tensors
(the U-Net is frozen during pruning: unet_2d_conditional.py:2118-2122)."""
from __future__ import annotations

import torch

from diffusion_pruning_b200.synthetic import split_arch
from unet_checks import build_pair, inputs, metrics

# gate-gradient tolerance (bf16 activations / gradients vs fp32 autograd): per gate family, the relative L2
# error of the gradient vector must stay below REL_TOL and its cosine above COS_TOL.
REL_TOL = 6e-2
COS_TOL = 0.995


def _loss(pred, taps, seed=5):
    g = torch.Generator().manual_seed(seed)
    total = 0.0
    for t in [pred] + list(taps):
        w = torch.randn(t.shape, generator=g).to(t.device)
        total = total + (t.float() * w).sum() / t[0].numel()
    return total


def check_train_grads(tiny=True, B=2, H=16, seed=9, use_taps=True):
    model, oracle = build_pair(tiny, beta_std=0.1)
    st = model.get_structure()
    dim = sum(w for ws in st["width"] for w in ws) + 14
    n_width = dim - 14
    g = torch.Generator().manual_seed(seed)
    arch = torch.rand(B, dim, generator=g) * 0.9 + 0.05
    sample, t, ctx = inputs(B, H, model.config["cross_attention_dim"])
    # oracle (CPU fp32 autograd)
    a_ref = arch.clone().requires_grad_(True)
    oracle.set_structure(split_arch(a_ref, st))
    pred_ref, taps_ref = oracle(sample, t, ctx, return_blocks=True)
    _loss(pred_ref, taps_ref if use_taps else []).backward()
    # product (CUDA): hooks capture the nine block outputs exactly like trainer.py:496-511
    acts = {}
    handles = []
    blocks = list(model.down_blocks) + [model.mid_block] + list(model.up_blocks)
    for i, blk in enumerate(blocks):
        def hook(mod, inp, out, i=i):
            acts[i] = out[0] if mod.kind == "down" else out
        handles.append(blk.register_forward_hook(hook))
    a_got = arch.clone().cuda().requires_grad_(True)
    model.set_structure(split_arch(a_got, st))
    pred = model(sample.cuda(), t.cuda(), ctx.cuda()).sample
    taps = [acts[i] for i in range(len(blocks))]
    _loss(pred, taps if use_taps else []).backward()
    torch.cuda.synchronize()
    from diffusion_pruning_b200 import kernels as K
    K.check_abort()
    for h in handles:
        h.remove()
    out = {"pred": metrics(pred.detach(), pred_ref.detach())}
    for i, (a, b) in enumerate(zip(taps, taps_ref)):
        out[f"tap{i}"] = metrics(a.detach(), b.detach())
    got, ref = a_got.grad.cpu(), a_ref.grad
    fam = {"width": (0, n_width), "depth": (n_width, dim)}
    for name, (lo, hi) in fam.items():
        gg, rr = got[:, lo:hi].flatten(), ref[:, lo:hi].flatten()
        rel = ((gg - rr).norm() / rr.norm().clamp_min(1e-12)).item()
        cos = torch.nn.functional.cosine_similarity(gg, rr, dim=0).item()
        out[f"grad_{name}"] = (rel, cos)
    return out


def assert_train(out):
    from unet_checks import COS_TOL as FWD_COS, MAX_ABS_TOL
    for k, (a, b) in out.items():
        if k.startswith("grad_"):
            assert a <= REL_TOL and b >= COS_TOL, f"{k}: relative L2 error {a:.4g}, cosine {b:.6f}"
        else:
            assert a <= 2 * MAX_ABS_TOL and b >= FWD_COS - 5e-4, f"{k}: max_abs {a:.4g} cosine {b:.6f}"
