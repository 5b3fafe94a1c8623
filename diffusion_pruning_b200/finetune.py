"""The fine-tune step of one static expert on the B200 path: a restatement of `FineTuner.step`
(pdm/training/trainer.py:1683-1765) from the point where the batch has been encoded (VAE / CLIP are out of scope,
SURVEY 2.1), i.e. it consumes noisy latents, timesteps, the prediction target and text-encoder states and returns the
reference's (loss, diff_loss, distillation_loss, block_loss). The student is a gated U-Net fixed to one architecture
code with `enable_weight_training()`: its forward / backward run on the tape engine (`train.UNetFineTuneFunction`) and
produce gradients for every U-Net parameter through the sm_100a kernels (dgrad, K8 wgrad, norm-affine, attention
backward); the teacher is the dense U-Net (all-ones gates) under `torch.no_grad()`; the losses are the K6 kernels.

Semantics: a `UNet2DConditionModelGated` student trains with gated semantics (a gated-off GroupNorm group still feeds
silu(beta) into conv2); a `UNet2DConditionModelPruned` student -- what the reference's FineTuner
trains (trainer.py:1452-1462) -- trains with prune() semantics: on the dense weights that is norm2.bias masked by the
expert's hard gate (and a masked bias gradient), checked against the autograd of the physically sliced oracle. Either
way the gated-off channels are still computed (dense GEMMs) and their gradients are exactly zero, so an optimizer step
never moves them; compacting the training GEMMs like the hard-gate forward is the next step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch

from .losses import block_mse, min_snr_weights, prediction_losses
from .pruning_step import BlockTaps, alphas_cumprod


@dataclass
class FinetuneLossConfig:
    """configs/finetuning/sd-2-1_cc3m.yaml:86-96."""
    snr_gamma: Optional[float] = 5.0
    diffusion_weight: float = 0.01
    distillation_weight: float = 0.5
    block_weight: float = 0.5
    prediction_type: str = "v_prediction"


def finetune_step(unet, teacher, batch: Dict[str, torch.Tensor], cfg: FinetuneLossConfig, taps: BlockTaps,
                  teacher_taps: BlockTaps, acp: Optional[torch.Tensor] = None) -> Dict[str, Any]:
    """batch keys: noisy_latents [B,4,H,W], timesteps [B] int64, target [B,4,H,W], encoder_hidden_states [B,77,1024]."""
    noisy, timesteps, target = batch["noisy_latents"], batch["timesteps"], batch["target"]
    enc = batch["encoder_hidden_states"]
    with torch.no_grad():                                              # trainer.py:1727-1728
        full_pred = teacher(noisy, timesteps, enc).sample.detach()
        teacher_acts = dict(teacher_taps.acts)
    model_pred = unet(noisy, timesteps, enc).sample                    # :1730
    student_acts = dict(taps.acts)
    w = None
    if cfg.snr_gamma is not None:                                      # :1731-1749
        acp = alphas_cumprod() if acp is None else acp
        w = min_snr_weights(acp, timesteps, cfg.snr_gamma, cfg.prediction_type == "v_prediction")
    diff, distill = prediction_losses(model_pred, target, full_pred, w)
    loss = cfg.diffusion_weight * diff                                 # :1751-1752
    block = torch.zeros((), device=noisy.device)
    if cfg.block_weight > 0:                                           # :1754-1759
        for k in student_acts:
            block = block + block_mse(student_acts[k], teacher_acts[k])
        block = block / len(student_acts)
        loss = loss + cfg.block_weight * block
    loss = loss + cfg.distillation_weight * distill                    # :1761-1762
    return {"loss": loss, "diff_loss": diff.detach(), "distillation_loss": distill, "block_loss": block}
