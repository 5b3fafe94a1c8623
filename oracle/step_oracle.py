"""CPU fp32 oracle of the pruning train step (pdm/training/trainer.py:1092-1254, from the encoded batch on)
-- TEST INFRASTRUCTURE, not product code. Composed from oracle/router_oracle.py (pinned to reference-generated
goldens) and oracle/unet_oracle.py; plain autograd. The Gumbel uniforms are drawn from the global CPU
generator in the reference's order: codebook (quantizer.py:141), z inside the OT routing (:325), the
architecture vector (trainer.py:1132), the codebook again (trainer.py:1140)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import router_oracle as R


def alphas_cumprod(n=1000, b0=0.00085, b1=0.012):
    betas = torch.linspace(b0 ** 0.5, b1 ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def add_noise(latents: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor, acp: torch.Tensor) -> torch.Tensor:
    """diffusers 0.23.1 DDIMScheduler.add_noise as called at trainer.py:1121-1123."""
    sa = (acp ** 0.5)[timesteps].flatten()
    sb = ((1 - acp) ** 0.5)[timesteps].flatten()
    while sa.dim() < latents.dim():
        sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
    return sa * latents + sb * noise


def get_velocity(latents: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor, acp: torch.Tensor) -> torch.Tensor:
    """diffusers 0.23.1 DDIMScheduler.get_velocity as called at trainer.py:1181."""
    sa = (acp ** 0.5)[timesteps].flatten()
    sb = ((1 - acp) ** 0.5)[timesteps].flatten()
    while sa.dim() < latents.dim():
        sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
    return sa * noise - sb * latents


def min_snr_weights(acp: torch.Tensor, timesteps: torch.Tensor, snr_gamma: float, v_prediction: bool) -> torch.Tensor:
    """trainer.py:1201-1212 with compute_snr of pdm/utils/metric_utils.py:3-26."""
    snr = ((acp ** 0.5)[timesteps] / ((1 - acp) ** 0.5)[timesteps]) ** 2
    if v_prediction:
        snr = snr + 1
    return torch.stack([snr, snr_gamma * torch.ones_like(timesteps)], dim=1).min(dim=1)[0] / snr


def split(arch: torch.Tensor, layout: R.ArchLayout) -> Dict[str, list]:
    """hypernet.py:86-101."""
    ws = layout.width_starts
    return {"width": [arch[:, ws[i]:ws[i + 1]] for i in range(len(layout.width_list))],
            "depth": [arch[:, layout.n_width + i] for i in range(layout.n_depth)]}


def pruning_step(unet, hyper_w: torch.Tensor, hyper_b: torch.Tensor, codebook: torch.Tensor, layout: R.ArchLayout,
                 batch: Dict[str, torch.Tensor], cfg, p_actual: float, temperature=0.4, base=3.0, ddp=None):
    """unet: GatedUNetOracle with count_macs() done and `ones_prunable` = cur_prunable at all-ones gates.

    ddp (data-parallel emulation in ONE process, for the 2-rank NCCL test): "collect" returns this rank's router
    tensors before the U-Net runs; a dict(idx, text_all, arch_all, rank) replaces the local Sinkhorn assignment by this
    rank's slice of the GLOBAL one (quantizer.py:278-300: marginals all-reduced over ranks) and evaluates the contrastive
    loss on the all-gathered batch in which only the local rows carry gradient (trainer.py:1153-1160)."""
    noisy, timesteps, target = batch["noisy_latents"], batch["timesteps"], batch["target"]
    enc, text = batch["encoder_hidden_states"], batch["mpnet_embeddings"]
    B, K = text.shape[0], codebook.shape[0]
    arch = F.linear(text, hyper_w, hyper_b)                                          # trainer.py:1129
    # quantizer.forward, train mode (quantizer.py:140-151)
    codes_gs = R.gumbel_sigmoid_trick(codebook, R.draw_uniforms(layout, K, False), layout, temperature, base)
    z_gs = R.gumbel_sigmoid_trick(arch.detach(), R.draw_uniforms(layout, B, False), layout, temperature, base)
    if isinstance(ddp, dict):
        idx = ddp["idx"]
    else:
        idx, _, _ = R.ot_indices(z_gs, codes_gs.detach(), layout)
    arch_q = codes_gs[idx]
    arch_gs = R.gumbel_sigmoid_trick(arch, R.draw_uniforms(layout, B, False), layout, temperature, base)  # :1132
    arch_norm = R.width_depth_normalize(arch_gs, layout)                             # :1138
    _ = R.draw_uniforms(layout, K, False)                                            # :1140 (similarity, no grad)
    if ddp == "collect":
        return {"z_gs": z_gs.detach(), "codes_gs": codes_gs.detach(), "arch_norm": arch_norm.detach()}
    if isinstance(ddp, dict):
        r, Bl = ddp["rank"], text.shape[0]
        arch_all = torch.cat([ddp["arch_all"][:r * Bl], arch_norm, ddp["arch_all"][(r + 1) * Bl:]], 0)
        c_loss = R.contrastive_loss(ddp["text_all"], arch_all, cfg.arch_vector_temperature,
                                    cfg.prompt_embedding_temperature)
    else:
        c_loss = R.contrastive_loss(text, arch_norm, cfg.arch_vector_temperature, cfg.prompt_embedding_temperature)
    with torch.no_grad():                                                            # :1185-1190
        unet.set_structure(split(torch.ones_like(arch_gs), layout))
        full_pred, t_taps = unet(noisy, timesteps, enc, return_blocks=True)
    unet.set_structure(split(arch_q, layout))                                        # :1192-1195
    pred, s_taps = unet(noisy, timesteps, enc, return_blocks=True)
    acp = alphas_cumprod()
    snr = ((acp ** 0.5)[timesteps] / ((1 - acp) ** 0.5)[timesteps]) ** 2
    if cfg.prediction_type == "v_prediction":
        snr = snr + 1
    w = torch.stack([snr, cfg.snr_gamma * torch.ones_like(timesteps)], dim=1).min(dim=1)[0] / snr
    loss = F.mse_loss(pred, target, reduction="none")
    loss = (loss.mean(dim=[1, 2, 3]) * w).mean()
    distill = F.mse_loss(pred, full_pred)
    block = sum(F.mse_loss(a, b.detach()) for a, b in zip(s_taps, t_taps)) / len(s_taps)
    macs = unet.calc_macs()
    ratios = macs["cur_prunable_macs"] / unet.ones_prunable
    r_loss = R.resource_loss(ratios.mean(), p_actual, cfg.resource_type)
    max_loss = 1.0 - torch.max(ratios)
    std_loss = -torch.std(ratios)
    total = (cfg.diffusion_weight * loss + cfg.resource_weight * r_loss + cfg.contrastive_weight * c_loss +
             cfg.distillation_weight * distill + cfg.block_weight * block + cfg.std_weight * std_loss +
             cfg.max_weight * max_loss)
    return {"loss": total, "diff_loss": loss.detach(), "distillation_loss": distill, "block_loss": block,
            "contrastive_loss": c_loss, "resource_loss": r_loss, "resource_ratio": ratios.mean().detach(),
            "idx": idx, "arch_q": arch_q}


def finetune_step(student, teacher, batch: Dict[str, torch.Tensor], cfg):
    """FineTuner.step (pdm/training/trainer.py:1683-1765) from the encoded batch on: `student` = GatedUNetOracle fixed
    to one code, `teacher` = the dense oracle (all-ones gates)."""
    noisy, timesteps, target = batch["noisy_latents"], batch["timesteps"], batch["target"]
    enc = batch["encoder_hidden_states"]
    with torch.no_grad():
        full_pred, t_taps = teacher(noisy, timesteps, enc, return_blocks=True)
    pred, s_taps = student(noisy, timesteps, enc, return_blocks=True)
    acp = alphas_cumprod()
    w = min_snr_weights(acp, timesteps, cfg.snr_gamma, cfg.prediction_type == "v_prediction")
    diff = (F.mse_loss(pred, target, reduction="none").mean(dim=[1, 2, 3]) * w).mean()
    loss = cfg.diffusion_weight * diff
    block = sum(F.mse_loss(a, b.detach()) for a, b in zip(s_taps, t_taps)) / len(s_taps)
    loss = loss + cfg.block_weight * block
    distill = F.mse_loss(pred, full_pred)
    loss = loss + cfg.distillation_weight * distill
    return {"loss": loss, "diff_loss": diff.detach(), "distillation_loss": distill, "block_loss": block}
