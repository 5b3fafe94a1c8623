"""Replays the reference's CALL SITES of the hot path against the drop-in classes, statement by statement, so a
maintainer switching `pdm.models` for `diffusion_pruning_b200` finds every attribute / call the callers make:

  Pruner.count_macs          pdm/training/trainer.py:1257-1296  (count_ops_and_params replaced by a plain forward)
  pipeline routing + loop    pdm/pipelines/pruning_pipelines.py:746-759, :772, :790-824
  FineTuner.init_models      pdm/training/trainer.py:1440-1462  (config.in_channels / sample_size, arch_vector kwarg)

The scheduler arithmetic (diffusers DDIMScheduler, v-prediction, eta 0) is restated inline as in oracle/sampling_oracle.py.
"""
from __future__ import annotations

import torch

from diffusion_pruning_b200 import HyperStructure, StructureVectorQuantizer, UNet2DConditionModelGated
from diffusion_pruning_b200.synthetic import DEPTH_ORDER, synthetic_codes

TINY = dict(block_out_channels=(64, 128, 256, 256), attention_head_dim=(1, 2, 4, 4), cross_attention_dim=128,
            sample_size=16)


def build(device):
    torch.manual_seed(0)
    unet = UNet2DConditionModelGated(**TINY).to(device).eval()
    st = unet.get_structure()
    hyper_net = HyperStructure(structure=st, input_dim=32, wn_flag=False, linear_bias=True).to(device).eval()
    quantizer = StructureVectorQuantizer(n_e=8, structure=st, beta=0.25, temperature=0.4, base=3,
                                         depth_order=list(DEPTH_ORDER), non_zero_width=True,
                                         resource_aware_normalization=False, optimal_transport=True).to(device).eval()
    codes = synthetic_codes(st, 8).float()
    quantizer.embedding_gs.data = (codes * 0.9 + 0.05).to(device)
    return unet, hyper_net, quantizer


def replay_count_macs(unet, hyper_net, quantizer, device, H=16):
    """trainer.py:1257-1296 with `count_ops_and_params(self.unet, {...})` (:1272) replaced by the forward it performs."""
    arch_vecs_separated = hyper_net.transform_structure_vector(
        torch.ones((1, quantizer.vq_embed_dim), device=device))                                   # :1262-1263
    unet.set_structure(arch_vecs_separated)                                                        # :1265
    latents = torch.randn(1, unet.config.in_channels, H, H, device=device)                         # :1267
    timesteps = torch.randint(0, 1000, (1,), device=device).long()                                 # :1268-1269
    encoder_hidden_states = torch.randn(1, 77, unet.config.cross_attention_dim, device=device)     # :1270
    with torch.no_grad():
        unet(sample=latents, timestep=timesteps, encoder_hidden_states=encoder_hidden_states)      # :1272 (hooked forward)
    sanity_macs_dict = unet.calc_macs()                                                            # :1281
    prunable_macs_list = [[e / sanity_macs_dict['prunable_macs'] for e in elem] for elem in
                          unet.get_prunable_macs()]                                                # :1282-1283
    unet.prunable_macs_list = prunable_macs_list                                                   # :1285
    unet.resource_info_dict = sanity_macs_dict                                                     # :1286
    quantizer.set_prunable_macs_template(prunable_macs_list)                                       # :1288
    out = {}
    for k, v in sanity_macs_dict.items():                                                          # :1290-1296
        out[k] = v.item() if isinstance(v, torch.Tensor) else v
    return out


def replay_pipeline(unet, hyper_net, quantizer, device, n_prompts=3, steps=3, guidance_scale=7.5, H=None):
    """pruning_pipelines.py:746-824 from the encoded prompt on. Returns (latents, indices, resource_ratios)."""
    do_classifier_free_guidance = guidance_scale > 1.0
    g = torch.Generator().manual_seed(5)
    H = H or unet.config.sample_size                                                              # :708-709
    prompt_cond = torch.randn(n_prompts, 77, unet.config.cross_attention_dim, generator=g).to(device)
    negative = torch.randn(1, 77, unet.config.cross_attention_dim, generator=g).to(device).expand(n_prompts, -1, -1)
    hyper_net_input = torch.randn(n_prompts, 32, generator=g).to(device)
    structure_vector = hyper_net(hyper_net_input)                                                  # :746-749
    structure_vector_quantized, (_, _, min_encoding_indices) = quantizer(structure_vector)         # :751
    structure_vector = quantizer.gumbel_sigmoid_trick(structure_vector)                            # :753
    arch_vectors_separated = hyper_net.transform_structure_vector(structure_vector_quantized)      # :757
    unet.set_structure(arch_vectors_separated)                                                     # :759
    prompt_embeds = torch.cat([negative, prompt_cond]) if do_classifier_free_guidance else prompt_cond  # :764-765
    num_channels_latents = unet.config.in_channels                                                 # :772
    latents = torch.randn(n_prompts, num_channels_latents, H, H, generator=g).to(device)
    # DDIM, SD-2.1 scheduler config (leading spacing, steps_offset 1, v-prediction, eta 0)
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    acp = torch.cumprod(1.0 - betas, 0)
    ratio = 1000 // steps
    timesteps = (torch.arange(0, steps) * ratio).flip(0) + 1
    trace = []
    with torch.no_grad():
        for i, t in enumerate(timesteps):                                                          # :790
            latent_model_input = torch.cat([latents] * 2) if do_classifier_free_guidance else latents  # :792
            noise_pred = unet(latent_model_input, t.to(device), encoder_hidden_states=prompt_embeds,
                              cross_attention_kwargs=None, return_dict=False)[0]                   # :796-802
            trace.append(noise_pred.clone())  # raw [uncond; cond] prediction of this step (test hook, not in the reference)
            if do_classifier_free_guidance:                                                        # :805-807
                noise_pred_uncond, noise_pred_text = noise_pred.chunk(2)
                noise_pred = noise_pred_uncond + guidance_scale * (noise_pred_text - noise_pred_uncond)
            a_t = acp[int(t)]
            prev = int(t) - ratio
            a_p = acp[prev] if prev >= 0 else acp[0]
            x0 = a_t.sqrt() * latents - (1 - a_t).sqrt() * noise_pred
            eps = a_t.sqrt() * noise_pred + (1 - a_t).sqrt() * latents
            latents = a_p.sqrt() * x0 + (1 - a_p).sqrt() * eps                                      # :814
    macs_dict = unet.calc_macs()                                                                   # :822
    resource_ratios = macs_dict['cur_prunable_macs'] / (unet.resource_info_dict['cur_prunable_macs'].squeeze())  # :823-824
    return latents, min_encoding_indices, resource_ratios, structure_vector_quantized, trace
