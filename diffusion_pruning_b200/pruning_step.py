"""The pruning train step of APTP on the B200 hot path: a restatement of `Pruner.step`
(pdm/training/trainer.py:1092-1254) from the point where the batch has been encoded (VAE / CLIP / MPNet
are out of scope: SURVEY section 2.1 rows 9-11), i.e. it consumes noisy latents, timesteps, the
prediction target, text-encoder states and MPNet prompt embeddings, and produces the reference's loss
tuple. Everything numerical runs in the sm_100a kernels through the drop-in modules:

  hyper_net -> quantizer (Gumbel-sigmoid codes, Sinkhorn OT routing, NCCL-reduced marginals when sharded)
  teacher U-Net forward (all-ones gates, no_grad)      trainer.py:1185-1190
  student U-Net forward (soft gates, one autograd node) trainer.py:1192-1195
  DDPM (min-SNR) + distillation + block + resource + contrastive + std/max losses  trainer.py:1197-1249

Only the scalar loss algebra on top of the kernel outputs uses PyTorch ops.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .losses import block_mse, min_snr_weights, prediction_losses


@dataclass
class PruningLossConfig:
    """configs/pruning/sd-2-1_cc3m.yaml:86-111."""
    snr_gamma: Optional[float] = 5.0
    diffusion_weight: float = 1.0
    resource_type: str = "log"
    resource_weight: float = 2.0
    pruning_target: float = 0.6
    arch_vector_temperature: float = 0.03
    prompt_embedding_temperature: float = 0.03
    contrastive_weight: float = 100.0
    distillation_weight: float = 0.2
    block_weight: float = 0.2
    std_weight: float = 0.1
    max_weight: float = 0.1
    prediction_type: str = "v_prediction"


def alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012):
    """SD-2.1 scheduler: scaled-linear betas (diffusers DDIMScheduler / DDPMScheduler config)."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def compute_snr(acp: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
    """pdm/utils/metric_utils.py:3-26."""
    a = (acp ** 0.5).to(timesteps.device)[timesteps].float()
    s = ((1.0 - acp) ** 0.5).to(timesteps.device)[timesteps].float()
    return (a / s) ** 2


def contrastive_loss(prompt_embeddings, arch_vectors, t_arch: float, t_prompt: float):
    """pdm/losses/contrastive_loss.py:11-22: the fused K9 kernels on CUDA tensors (csrc/smallops.cu); the torch-op
    form below only serves CPU tensors (host-side checks of the loss algebra in the GPU-less container)."""
    if arch_vectors.is_cuda:
        from .losses import contrastive_loss as _cl
        return _cl(prompt_embeddings, arch_vectors, t_arch, t_prompt)
    a = arch_vectors / arch_vectors.norm(dim=1, keepdim=True)
    p = prompt_embeddings / prompt_embeddings.norm(dim=1, keepdim=True)
    sa = F.softmax((a @ a.T) / t_arch, dim=-1)
    sp = F.softmax((p @ p.T) / t_prompt, dim=-1)
    return F.binary_cross_entropy(sa.T, sp.T, reduction="mean"), sa.detach()


def resource_loss(ratio: torch.Tensor, p: float, loss_type: str = "log"):
    """pdm/losses/resource_loss.py:12-23 without the host sync of `if ratio > p`: |log(ratio / p)|."""
    if loss_type == "log":
        return torch.abs(torch.log(ratio / p))
    if loss_type == "mae":
        return torch.abs(ratio - p)
    return (ratio - p) ** 2


def actual_pruning_target(unet, p: float) -> float:
    """Pruner.update_pruning_target (trainer.py:1299-1306)."""
    ri = unet.resource_info_dict
    return float(1 - (1 - p) * ri["total_macs"] / ri["cur_prunable_macs"])


class BlockTaps:
    """The trainer's forward hooks on down_blocks[i] / mid_block / up_blocks[i] (trainer.py:496-511)."""

    def __init__(self, unet):
        self.acts: Dict[str, torch.Tensor] = {}
        self.handles = []
        for i, blk in enumerate(unet.down_blocks):
            self.handles.append(blk.register_forward_hook(self._hook("d%d" % i, True)))
        self.handles.append(unet.mid_block.register_forward_hook(self._hook("m", False)))
        for i, blk in enumerate(unet.up_blocks):
            self.handles.append(blk.register_forward_hook(self._hook("u%d" % i, False)))

    def _hook(self, name, residuals_present):
        def hook(module, inp, out):
            self.acts[name] = out[0] if residuals_present else out
        return hook

    def remove(self):
        for h in self.handles:
            h.remove()


def pruning_step(unet, hyper_net, quantizer, batch: Dict[str, torch.Tensor], cfg: PruningLossConfig, taps: BlockTaps,
                 p_actual: float, acp: Optional[torch.Tensor] = None, pretrain: bool = False) -> Dict[str, Any]:
    """One `Pruner.step` from the encoded batch on. batch keys: noisy_latents [B,4,H,W], timesteps [B] int64,
    target [B,4,H,W], encoder_hidden_states [B,77,1024], mpnet_embeddings [B,768]."""
    noisy, timesteps, target = batch["noisy_latents"], batch["timesteps"], batch["target"]
    enc, text = batch["encoder_hidden_states"], batch["mpnet_embeddings"]
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0

    # The teacher forward (trainer.py:1185-1190) depends on the batch only, not on the architecture vector, so it is
    # ENQUEUED FIRST: its ~500 kernels (one CUDA-graph replay) keep the GPU busy while the host walks through the
    # hyper-network / quantizer / contrastive-loss code below, whose small kernels and host<->device handshakes would
    # otherwise leave the GPU idle between two steps. Results are identical (the teacher consumes no random numbers).
    with torch.no_grad():
        full = hyper_net.transform_structure_vector(
            torch.ones(text.shape[0], quantizer.vq_embed_dim, device=text.device, dtype=torch.float32))
        unet.set_structure(full)
        full_pred = unet(noisy, timesteps, enc).sample.detach()
        teacher_acts = dict(taps.acts)

    arch_vector = hyper_net(text)                                      # trainer.py:1129
    arch_vector_quantized, _ = quantizer(arch_vector)                  # :1130
    arch_vector = quantizer.gumbel_sigmoid_trick(arch_vector)          # :1132
    arch_norm = quantizer.width_depth_normalize(arch_vector)           # :1138
    with torch.no_grad():                                              # :1139-1160
        q_emb = quantizer.get_codebook_entry_gumbel_sigmoid(torch.arange(quantizer.n_e, device=text.device),
                                                            hard=True).detach()
        q_emb = q_emb / q_emb.norm(dim=-1, keepdim=True)
        q_sim = q_emb @ q_emb.t()
        if world > 1:
            # ONE collective for both gathers of trainer.py:1153-1154: [B, 768 + 1620] fp32 rows, gathered in rank order
            packed = torch.cat([text.to(torch.float32), arch_norm.detach().to(torch.float32)], dim=1).contiguous()
            gathered = torch.empty(world * packed.shape[0], packed.shape[1], device=packed.device, dtype=packed.dtype)
            dist.all_gather_into_tensor(gathered, packed)
    if world > 1:
        Bl, nt = text.shape[0], text.shape[1]
        text_all = gathered[:, :nt].to(text.dtype)
        # the local slice keeps its gradient (:1159-1160); the other ranks' rows are constants
        arch_all = torch.cat([gathered[:rank * Bl, nt:], arch_norm, gathered[(rank + 1) * Bl:, nt:]], 0)
    else:
        text_all, arch_all = text, arch_norm
    separated = hyper_net.transform_structure_vector(arch_vector if pretrain else arch_vector_quantized)
    c_loss, _ = contrastive_loss(text_all, arch_all, cfg.arch_vector_temperature, cfg.prompt_embedding_temperature)

    unet.set_structure(separated)                                      # student, :1192-1195
    model_pred = unet(noisy, timesteps, enc).sample
    student_acts = dict(taps.acts)

    # :1197-1225 on the K6 kernels: one pass over (pred, target, teacher pred) for the DDPM + distillation losses,
    # one pass per hooked block output for the block losses (bf16 NHWC in place, gradient in the same layout)
    w = None
    if cfg.snr_gamma is not None:                                      # :1197-1216
        acp = alphas_cumprod() if acp is None else acp
        w = min_snr_weights(acp, timesteps, cfg.snr_gamma, cfg.prediction_type == "v_prediction")
    loss, distill = prediction_losses(model_pred, target, full_pred, w)  # :1197-1218
    block = torch.zeros((), device=noisy.device)
    for k in student_acts:                                             # :1220-1225
        block = block + block_mse(student_acts[k], teacher_acts[k])
    block = block / len(student_acts)

    macs = unet.calc_macs()                                            # :1227-1249
    ratios = macs["cur_prunable_macs"] / unet.resource_info_dict["cur_prunable_macs"].squeeze()
    r_loss = resource_loss(ratios.mean(), p_actual, cfg.resource_type)
    max_loss = 1.0 - torch.max(ratios)
    std_loss = -torch.std(ratios)
    diff_loss = loss.detach().clone()
    total = (cfg.diffusion_weight * loss + cfg.resource_weight * r_loss + cfg.contrastive_weight * c_loss +
             cfg.distillation_weight * distill + cfg.block_weight * block + cfg.std_weight * std_loss +
             cfg.max_weight * max_loss)
    return {"loss": total, "diff_loss": diff_loss, "distillation_loss": distill, "block_loss": block,
            "contrastive_loss": c_loss, "resource_loss": r_loss, "resource_ratio": ratios.mean().detach(),
            "macs": macs, "arch_vector_quantized": arch_vector_quantized, "codebook_similarity": q_sim,
            "std_loss": std_loss, "max_loss": max_loss}
