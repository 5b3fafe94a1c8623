"""K6 loss kernels (add_noise / velocity, block MSE, prediction losses) vs the CPU oracle / fp32 autograd."""
import torch

from oracle import step_oracle as SO


def check_add_noise(B=5, C=4, H=24, seed=3):
    from diffusion_pruning_b200 import losses as L
    g = torch.Generator().manual_seed(seed)
    x, n = torch.randn(B, C, H, H, generator=g), torch.randn(B, C, H, H, generator=g)
    t = torch.tensor([0, 999, 500, 17, 981][:B])
    acp = SO.alphas_cumprod()
    tabs = L.NoiseTables(acp, "cuda")
    out = {}
    for vp in (True, False):
        noisy, target = L.add_noise_and_target(tabs, x.cuda(), n.cuda(), t.cuda(), v_prediction=vp)
        ref_noisy = SO.add_noise(x, n, t, acp)
        ref_target = SO.get_velocity(x, n, t, acp) if vp else n
        out[vp] = ((noisy.cpu() - ref_noisy).abs().max().item(), (target.cpu() - ref_target).abs().max().item())
    return out


def check_block_mse(B=3, C=48, H=10, W=6, pitch=64, seed=4):
    """Pitched channels-last bf16 views (what the engine hands to the hooks) and a plain NCHW tensor."""
    from diffusion_pruning_b200 import losses as L
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, ld in (("pitched", pitch), ("dense", C)):
        s_buf = torch.randn(B, H, W, ld, generator=g).to(torch.bfloat16).cuda()
        t_buf = (s_buf.float() + 0.3 * torch.randn(B, H, W, ld, generator=g).cuda()).to(torch.bfloat16)
        s = s_buf[..., :C].permute(0, 3, 1, 2).requires_grad_(True)
        t = t_buf[..., :C].permute(0, 3, 1, 2)
        loss = L.block_mse(s, t)
        (loss * 3.0).backward()
        s_ref = s.detach().float().cpu().requires_grad_(True)
        ref = torch.nn.functional.mse_loss(s_ref, t.float().cpu())
        (ref * 3.0).backward()
        gd = (s.grad.float().cpu() - s_ref.grad).abs().max().item() / s_ref.grad.abs().max().item()
        out[name] = (abs(loss.item() - ref.item()) / ref.item(), gd)
    s = torch.randn(2, 16, 4, 4, generator=g).to(torch.bfloat16).cuda().requires_grad_(True)  # NCHW-contiguous input
    t = torch.randn(2, 16, 4, 4, generator=g).to(torch.bfloat16).cuda()
    loss = L.block_mse(s, t)
    loss.backward()
    ref = torch.nn.functional.mse_loss(s.detach().float(), t.float())
    out["nchw"] = (abs(loss.item() - ref.item()) / ref.item(),
                   ((s.grad.float() - 2 * (s.detach().float() - t.float()) / s.numel()).abs().max() /
                    s.grad.float().abs().max()).item())
    return out


def check_pred_losses(B=4, H=16, seed=6):
    from diffusion_pruning_b200 import losses as L
    g = torch.Generator().manual_seed(seed)
    pred, target, teacher = (torch.randn(B, 4, H, H, generator=g) for _ in range(3))
    t = torch.tensor([981, 661, 341, 21][:B])
    acp = SO.alphas_cumprod()
    out = {}
    for gamma in (5.0, None):
        p_ref = pred.clone().requires_grad_(True)
        if gamma is None:
            l_ref = torch.nn.functional.mse_loss(p_ref, target)
            w_ref = None
        else:
            w_ref = SO.min_snr_weights(acp, t, gamma, True)
            l_ref = (torch.nn.functional.mse_loss(p_ref, target, reduction="none").mean(dim=[1, 2, 3]) * w_ref).mean()
        d_ref = torch.nn.functional.mse_loss(p_ref, teacher)
        (1.0 * l_ref + 0.5 * d_ref).backward()
        p = pred.clone().cuda().requires_grad_(True)
        w = None if gamma is None else L.min_snr_weights(acp, t.cuda(), gamma, True)
        l, d = L.prediction_losses(p, target.cuda(), teacher.cuda(), w)
        (1.0 * l + 0.5 * d).backward()
        werr = 0.0 if gamma is None else (w.cpu() - w_ref).abs().max().item()
        out[gamma] = (abs(l.item() - l_ref.item()) / l_ref.item(), abs(d.item() - d_ref.item()) / d_ref.item(),
                      ((p.grad.cpu() - p_ref.grad).abs().max() / p_ref.grad.abs().max()).item(), werr)
    return out


def check_macs_kernel(B=6, H=16, seed=8):
    """aptp_macs_ratio_fwd/_bwd vs (1) the same closed form in torch ops on the same CUDA gates and (2) the CPU
    oracle's tree walk (oracle/unet_oracle.py calc_macs): values and straight-through gate gradients."""
    from diffusion_pruning_b200 import macs as M
    from unet_checks import build_pair, split_arch
    model, oracle = build_pair(True)
    model.count_macs(H, H)
    oracle.count_macs(H, H)
    st = model.get_structure()
    dim = sum(w for ws in st["width"] for w in ws) + sum(1 for d in st["depth"] if d == [1])
    g = torch.Generator().manual_seed(seed)
    arch = torch.rand(B, dim, generator=g)
    arch[0] = 1.0                      # all kept
    arch[1, -14:] = 0.1                # every depth-gated sub-block dropped
    wts = torch.randn(B, 1, generator=g)
    res = {}
    grads = {}
    for name in ("kernel", "torch", "oracle"):
        dev = "cpu" if name == "oracle" else "cuda"
        a = arch.clone().to(dev).requires_grad_(True)
        if name == "oracle":
            oracle.set_structure(split_arch(a, st))
            d = oracle.calc_macs()
        else:
            model.set_structure(split_arch(a, st))
            d = model.calc_macs() if name == "kernel" else M._calc_macs_torch(model)
        (d["cur_prunable_macs"] * wts.to(dev)).sum().backward()
        res[name] = (d["cur_prunable_macs"].detach().cpu().double(), d["cur_total_macs"].detach().cpu().double(),
                     float(d["total_macs"]), float(d["prunable_macs"]))
        grads[name] = a.grad.cpu().double()
    out = {}
    for other in ("torch", "oracle"):
        k, o = res["kernel"], res[other]
        out[other] = (((k[0] - o[0]).abs() / o[0].abs().clamp_min(1.0)).max().item(),
                      ((k[1] - o[1]).abs() / o[1].abs()).max().item(),
                      abs(k[2] - o[2]) / o[2], abs(k[3] - o[3]) / o[3],
                      ((grads["kernel"] - grads[other]).abs().max() / grads[other].abs().max()).item())
    return out


def check_hypernet_product_vs_reference_golden():
    """The PRODUCT HyperStructure (fp32 K9 linear kernel, no cuBLAS) against tests/golden/hypernet.npz, which was produced by
    executing the reference's own hypernet.py: forward values, the split into gate tensors, and the backward vs autograd of
    the same product in torch ops."""
    import os
    import numpy as np
    from diffusion_pruning_b200 import HyperStructure
    from oracle.structure import sd21_gate_structure
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hypernet.npz"))
    structure = sd21_gate_structure()
    hn = HyperStructure(structure=structure, input_dim=32, wn_flag=False, linear_bias=True).cuda()
    W, b = torch.from_numpy(g["weight"]), torch.from_numpy(g["bias"])
    off = 0
    with torch.no_grad():
        for l in hn.mh_fc:
            n = l.weight.shape[0]
            l.weight.copy_(W[off:off + n])
            l.bias.copy_(b[off:off + n])
            off += n
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    y = hn(x)
    assert torch.allclose(y.detach().cpu(), torch.from_numpy(g["y"]), atol=2e-6, rtol=1e-6), \
        (y.detach().cpu() - torch.from_numpy(g["y"])).abs().max()
    sv = hn.transform_structure_vector(y.detach())
    assert len(sv["width"]) == int(g["n_width"]) and len(sv["depth"]) == int(g["n_depth"])
    assert torch.allclose(sv["width"][3].cpu(), torch.from_numpy(g["width_3"]), atol=2e-6)
    assert torch.allclose(sv["depth"][5].cpu(), torch.from_numpy(g["depth_5"]), atol=2e-6)
    # backward: dW / db / dx vs torch autograd of x W^T + b
    gy = torch.randn_like(y)
    y.backward(gy)
    Wt, bt = W.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    xt = x.detach().clone().requires_grad_(True)
    (torch.nn.functional.linear(xt, Wt, bt) * gy).sum().backward()
    gw = torch.cat([l.weight.grad for l in hn.mh_fc], 0)
    gb = torch.cat([l.bias.grad for l in hn.mh_fc], 0)
    assert torch.allclose(gw, Wt.grad, atol=1e-5, rtol=1e-5) and torch.allclose(gb, bt.grad, atol=1e-5, rtol=1e-5)
    assert torch.allclose(x.grad, xt.grad, atol=1e-5, rtol=1e-5)


def check_contrastive_kernel_vs_reference_golden():
    """Fused contrastive-loss kernels vs the value the reference's own ContrastiveLoss produced (golden) and vs torch
    autograd of the same formula for the gradient."""
    import os
    import numpy as np
    from diffusion_pruning_b200.losses import contrastive_loss
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses_gates.npz"))
    p = torch.from_numpy(g["prompt"]).cuda()
    a = torch.from_numpy(g["arch"]).cuda().requires_grad_(True)
    loss, sa = contrastive_loss(p, a, 0.03, 0.03)
    assert abs(float(loss) - float(g["contrastive"])) <= 1e-5 * max(1.0, abs(float(g["contrastive"]))), (float(loss), float(g["contrastive"]))
    loss.backward()
    a2 = a.detach().clone().requires_grad_(True)
    an = a2 / a2.norm(dim=1, keepdim=True)
    pn = p / p.norm(dim=1, keepdim=True)
    s_a = torch.softmax(an @ an.T / 0.03, dim=-1)
    s_p = torch.softmax(pn @ pn.T / 0.03, dim=-1)
    ref = torch.nn.functional.binary_cross_entropy(s_a.T, s_p.T)
    ref.backward()
    assert torch.allclose(sa, s_a.detach(), atol=1e-6)
    rel = ((a.grad - a2.grad).norm() / a2.grad.norm()).item()
    assert rel <= 1e-4, rel
    # a larger, all-gathered-size batch (8 ranks x 32)
    gen = torch.Generator().manual_seed(3)
    p = torch.randn(256, 768, generator=gen).cuda()
    a = torch.rand(256, 1620, generator=gen).cuda().requires_grad_(True)
    loss, _ = contrastive_loss(p, a, 0.03, 0.03)
    (loss * 100.0).backward()
    a2 = a.detach().clone().requires_grad_(True)
    an = a2 / a2.norm(dim=1, keepdim=True)
    pn = p / p.norm(dim=1, keepdim=True)
    ref = torch.nn.functional.binary_cross_entropy(torch.softmax(an @ an.T / 0.03, -1).T, torch.softmax(pn @ pn.T / 0.03, -1).T)
    (ref * 100.0).backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
    rel = ((a.grad - a2.grad).norm() / a2.grad.norm()).item()
    assert rel <= 1e-3, rel
