"""CPU fp32 oracle of the routed sampling loop (pdm/pipelines/pruning_pipelines.py:757-820 from the encoded
prompts on; diffusers DDIMScheduler with the SD-2.1 config restated from SURVEY Appendix B: scaled-linear
betas, steps_offset 1, "leading" spacing, set_alpha_to_one False, eta 0) -- TEST INFRASTRUCTURE."""
from __future__ import annotations

import torch

from .step_oracle import alphas_cumprod, split


def ddim_timesteps(n: int, n_train: int = 1000, offset: int = 1):
    ratio = n_train // n
    return [i * ratio + offset for i in range(n)][::-1]


@torch.no_grad()
def denoise(unet, layout, arch: torch.Tensor, latents, cond, uncond, steps=25, guidance=7.5, v_prediction=True):
    acp = alphas_cumprod()
    unet.set_structure(split(arch, layout))  # gates tile over the doubled batch (gates.py:18-19)
    emb = torch.cat([uncond, cond], 0)       # pruning_pipelines.py:764-765
    x = latents.clone()
    ratio = 1000 // steps
    for t in ddim_timesteps(steps):
        x2 = torch.cat([x] * 2)              # :792
        pred = unet(x2, torch.full((x2.shape[0],), t), emb)
        pu, pc = pred.chunk(2)
        m = pu + guidance * (pc - pu)        # :805-807
        a_t = acp[t]
        a_prev = acp[t - ratio] if t - ratio >= 0 else acp[0]
        if v_prediction:
            x0 = a_t.sqrt() * x - (1 - a_t).sqrt() * m
            eps = a_t.sqrt() * m + (1 - a_t).sqrt() * x
        else:
            eps = m
            x0 = (x - (1 - a_t).sqrt() * m) / a_t.sqrt()
        x = a_prev.sqrt() * x0 + (1 - a_prev).sqrt() * eps   # DDIMScheduler.step, eta = 0
    return x
