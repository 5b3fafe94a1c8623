"""Routed-sampling parity: the CUDA loop (gated U-Net on the doubled CFG batch + fused CFG/DDIM kernel) vs the CPU
fp32 oracle loop on identical weights, codes, latents and text embeddings."""
from __future__ import annotations

import torch

from diffusion_pruning_b200 import HyperStructure
from diffusion_pruning_b200 import kernels as K
from diffusion_pruning_b200 import sampling as S
from diffusion_pruning_b200.synthetic import DEPTH_ORDER, synthetic_codes
from oracle import router_oracle as R
from oracle import sampling_oracle as SO
from unet_checks import build_pair, metrics


def check_cfg_ddim_kernel(seed=0):
    g = torch.Generator().manual_seed(seed)
    n = 3 * 4 * 16 * 16
    pred = torch.randn(2 * n, generator=g).cuda()
    x = torch.randn(n, generator=g).cuda()
    out = torch.empty_like(x)
    acp = SO.alphas_cumprod()
    for vpred in (True, False):
        K.cfg_ddim_step(pred, x, out, n, 7.5, float(acp[641]), float(acp[601]), vpred)
        pu, pc = pred[:n].cpu(), pred[n:].cpu()
        m = pu + 7.5 * (pc - pu)
        a, ap = acp[641], acp[601]
        xc = x.cpu()
        if vpred:
            x0, eps = a.sqrt() * xc - (1 - a).sqrt() * m, a.sqrt() * m + (1 - a).sqrt() * xc
        else:
            eps, x0 = m, (xc - (1 - a).sqrt() * m) / a.sqrt()
        ref = ap.sqrt() * x0 + (1 - ap).sqrt() * eps
        assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-5), f"cfg_ddim v_prediction={vpred}"
    assert S.ddim_timesteps(25) == SO.ddim_timesteps(25) and S.ddim_timesteps(25)[0] == 961


def check_sampling_loop(B=4, H=16, steps=4, code_ids=(0, 3, 3, 7), seed=3):
    model, oracle = build_pair(True, beta_std=0.1)
    st = model.get_structure()
    codes = synthetic_codes(st, 8)
    arch = codes[list(code_ids)]
    layout = R.ArchLayout(st, DEPTH_ORDER)
    hyper = HyperStructure(structure=st, input_dim=16, wn_flag=False, linear_bias=True)
    g = torch.Generator().manual_seed(seed)
    cd = model.config["cross_attention_dim"]
    lat = torch.randn(B, 4, H, H, generator=g)
    cond = torch.randn(B, 77, cd, generator=g)
    unc = torch.randn(B, 77, cd, generator=g)
    ref = SO.denoise(oracle, layout, arch.clone(), lat, cond, unc, steps=steps, guidance=7.5)
    got = S.denoise(model, hyper, arch.clone().cuda(), lat.cuda(), cond.cuda(), unc.cuda(), num_inference_steps=steps,
                    guidance_scale=7.5)
    torch.cuda.synchronize()
    K.check_abort()
    return metrics(got, ref)
