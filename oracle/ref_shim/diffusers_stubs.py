"""Stand-ins for the diffusers 0.23.1 classes the reference SUBCLASSES -- TEST INFRASTRUCTURE (golden generation only).

diffusers 0.23.1 (env.yaml:114) is neither vendored under /root/reference nor installable offline. The reference's
gating arithmetic, however, is written out in the reference's OWN files: every `forward` / `prune()` body cited in
SURVEY.md section 8(a) lives in pdm/models/unet/blocks.py and pdm/models/unet/unet_2d_conditional.py and only needs
its diffusers BASE CLASS to have built the sub-modules. This file provides exactly that: constructors with the
diffusers argument names that create the same nn.Modules under the same attribute names (= the state-dict keys), so
that the reference sources import and run IN PLACE and tests/golden/make_unet_goldens.py can record what they compute.

What stays restated here (and therefore is the only UNPINNED arithmetic of the U-Net oracle):
  * the container forwards the reference inherits untouched (CrossAttnDownBlock2D / DownBlock2D / CrossAttnUpBlock2D /
    UpBlock2D / UNetMidBlock2DCrossAttn.forward: the resnet -> attention loops, skip concatenation, samplers),
  * Transformer2DModel.forward for the width-only variant (same body as the reference's depth-gated override at
    blocks.py:1139-1355 minus the gate), Downsample2D / Upsample2D, Timesteps / TimestepEmbedding, GEGLU.gelu,
    FeedForward.forward (a sequential loop) and Attention.forward (dispatch to the processor).
Nothing under diffusion_pruning_b200/ imports this module, and it never travels to the GPU box as a dependency of
the product or of the GPU tests (they read the committed .npz fixtures).
"""
from __future__ import annotations

import inspect
import logging as _pylogging
import math
import os
import sys
import types
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------
# configuration_utils / modeling_utils
# ------------------------------------------------------------------------------------------------
class _Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def register_to_config(init):
    def wrapped(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = _Config({k: v for k, v in list(bound.arguments.items())[1:] if k not in ("args", "kwargs")})
        object.__setattr__(self, "_config", cfg)
        init(self, *args, **kwargs)
    wrapped.__wrapped__ = init
    return wrapped


class ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return self.__dict__.get("_config", _Config())

    def register_to_config(self, **kw):
        cfg = self.__dict__.get("_config")
        if cfg is None:
            cfg = _Config()
            object.__setattr__(self, "_config", cfg)
        cfg.update(kw)


class ModelMixin(nn.Module):
    _supports_gradient_checkpointing = True

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


# ------------------------------------------------------------------------------------------------
# activations / embeddings
# ------------------------------------------------------------------------------------------------
def get_activation(name: str) -> nn.Module:
    return {"swish": nn.SiLU, "silu": nn.SiLU, "gelu": nn.GELU, "relu": nn.ReLU, "mish": nn.Mish}[name.lower()]()


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def gelu(self, gate):
        return F.gelu(gate)

    def forward(self, hidden_states, scale: float = 1.0):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * self.gelu(gate)


class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float):
        super().__init__()
        self.num_channels, self.flip_sin_to_cos, self.downscale_freq_shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.downscale_freq_shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None, post_act_fn=None, cond_proj_dim=None):
        super().__init__()
        assert post_act_fn is None and cond_proj_dim is None
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.cond_proj = None
        self.act = get_activation(act_fn)
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim)
        self.post_act = None

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


# ------------------------------------------------------------------------------------------------
# resnet
# ------------------------------------------------------------------------------------------------
class Downsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, out_channels=None, padding=1, name="conv"):
        super().__init__()
        assert use_conv
        self.channels, self.out_channels, self.padding, self.name = channels, out_channels or channels, padding, name
        self.conv = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding)

    def forward(self, hidden_states, scale: float = 1.0):
        return self.conv(hidden_states)


class Upsample2D(nn.Module):
    def __init__(self, channels, use_conv=False, use_conv_transpose=False, out_channels=None, name="conv"):
        super().__init__()
        assert use_conv and not use_conv_transpose
        self.channels, self.out_channels, self.name = channels, out_channels or channels, name
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)

    def forward(self, hidden_states, output_size=None, scale: float = 1.0):
        if output_size is None:
            hidden_states = F.interpolate(hidden_states, scale_factor=2.0, mode="nearest")
        else:
            hidden_states = F.interpolate(hidden_states, size=output_size, mode="nearest")
        return self.conv(hidden_states)


class ResnetBlock2D(nn.Module):
    """Constructor only: the forward that runs is the reference's override (blocks.py:293-371, :482-584)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, temb_channels=512, groups=32,
                 groups_out=None, pre_norm=True, eps=1e-6, non_linearity="swish", skip_time_act=False,
                 time_embedding_norm="default", kernel=None, output_scale_factor=1.0, use_in_shortcut=None, up=False,
                 down=False, conv_shortcut_bias=True, conv_2d_out_channels=None):
        super().__init__()
        assert time_embedding_norm == "default" and not up and not down and kernel is None
        self.pre_norm = True
        self.in_channels = in_channels
        out_channels = in_channels if out_channels is None else out_channels
        self.out_channels = out_channels
        self.use_conv_shortcut = conv_shortcut
        self.up, self.down = up, down
        self.output_scale_factor = output_scale_factor
        self.time_embedding_norm = time_embedding_norm
        self.skip_time_act = skip_time_act
        groups_out = groups if groups_out is None else groups_out
        self.norm1 = nn.GroupNorm(num_groups=groups, num_channels=in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(num_groups=groups_out, num_channels=out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(dropout)
        conv_2d_out_channels = conv_2d_out_channels or out_channels
        self.conv2 = nn.Conv2d(out_channels, conv_2d_out_channels, kernel_size=3, stride=1, padding=1)
        self.nonlinearity = get_activation(non_linearity)
        self.upsample = self.downsample = None
        self.use_in_shortcut = self.in_channels != conv_2d_out_channels if use_in_shortcut is None else use_in_shortcut
        self.conv_shortcut = None
        if self.use_in_shortcut:
            self.conv_shortcut = nn.Conv2d(in_channels, conv_2d_out_channels, kernel_size=1, stride=1, padding=0,
                                           bias=conv_shortcut_bias)


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
class AttnProcessor2_0:
    def __init__(self):
        pass


class AttnProcessor(AttnProcessor2_0):
    pass


class AttnAddedKVProcessor(AttnProcessor2_0):
    pass


class Attention(nn.Module):
    """Constructor + processor dispatch; the arithmetic is HeadGatedAttnProcessor2.__call__ (blocks.py:194-280)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, cross_attention_norm=None, cross_attention_norm_num_groups=32,
                 added_kv_proj_dim=None, norm_num_groups=None, spatial_norm_dim=None, out_bias=True,
                 scale_qk=True, only_cross_attention=False, eps=1e-5, rescale_output_factor=1.0,
                 residual_connection=False, _from_deprecated_attn_block=False, processor=None):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention, self.upcast_softmax = upcast_attention, upcast_softmax
        self.rescale_output_factor, self.residual_connection = rescale_output_factor, residual_connection
        self.dropout = dropout
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        self.heads = heads
        self.sliceable_head_dim = heads
        self.added_kv_proj_dim, self.only_cross_attention = added_kv_proj_dim, only_cross_attention
        self.group_norm = None
        self.spatial_norm = None
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])
        self.set_processor(processor if processor is not None else AttnProcessor2_0())

    def set_processor(self, processor, _remove_lora=False):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **cross_attention_kwargs)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        assert activation_fn == "geglu"
        self.net = nn.ModuleList([GEGLU(dim, inner_dim), nn.Dropout(dropout), nn.Linear(inner_dim, dim_out)])
        if final_dropout:
            self.net.append(nn.Dropout(dropout))

    def forward(self, hidden_states, scale: float = 1.0):
        for module in self.net:
            hidden_states = module(hidden_states)
        return hidden_states


class BasicTransformerBlock(nn.Module):
    """Constructor only: the forward that runs is blocks.py:763-851."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, dropout=0.0, cross_attention_dim=None,
                 activation_fn="geglu", num_embeds_ada_norm=None, attention_bias=False, only_cross_attention=False,
                 double_self_attention=False, upcast_attention=False, norm_elementwise_affine=True,
                 norm_type="layer_norm", final_dropout=False, attention_type="default"):
        super().__init__()
        assert norm_type == "layer_norm" and num_embeds_ada_norm is None and attention_type == "default"
        self.only_cross_attention = only_cross_attention
        self.use_ada_layer_norm_zero = False
        self.use_ada_layer_norm = False
        self.norm1 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
        self.attn1 = Attention(query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                               bias=attention_bias, cross_attention_dim=cross_attention_dim if only_cross_attention else None,
                               upcast_attention=upcast_attention)
        if cross_attention_dim is not None or double_self_attention:
            self.norm2 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
            self.attn2 = Attention(query_dim=dim, cross_attention_dim=cross_attention_dim if not double_self_attention else None,
                                   heads=num_attention_heads, dim_head=attention_head_dim, dropout=dropout,
                                   bias=attention_bias, upcast_attention=upcast_attention)
        else:
            self.norm2 = None
            self.attn2 = None
        self.norm3 = nn.LayerNorm(dim, elementwise_affine=norm_elementwise_affine)
        self.ff = FeedForward(dim, dropout=dropout, activation_fn=activation_fn, final_dropout=final_dropout)
        self._chunk_size = None
        self._chunk_dim = 0


@dataclass
class Transformer2DModelOutput:
    sample: torch.FloatTensor


class Transformer2DModel(ModelMixin, ConfigMixin):
    def __init__(self, num_attention_heads=16, attention_head_dim=88, in_channels=None, out_channels=None, num_layers=1,
                 dropout=0.0, norm_num_groups=32, cross_attention_dim=None, attention_bias=False, sample_size=None,
                 num_vector_embeds=None, patch_size=None, activation_fn="geglu", num_embeds_ada_norm=None,
                 use_linear_projection=False, only_cross_attention=False, double_self_attention=False,
                 upcast_attention=False, norm_type="layer_norm", norm_elementwise_affine=True, attention_type="default"):
        super().__init__()
        assert in_channels is not None and patch_size is None and num_vector_embeds is None
        self.use_linear_projection = use_linear_projection
        self.num_attention_heads, self.attention_head_dim = num_attention_heads, attention_head_dim
        inner_dim = num_attention_heads * attention_head_dim
        self.is_input_continuous, self.is_input_vectorized, self.is_input_patches = True, False, False
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(num_groups=norm_num_groups, num_channels=in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner_dim)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList([])  # replaced by the reference's gated blocks (blocks.py:1015-1037)
        self.out_channels = in_channels if out_channels is None else out_channels
        if use_linear_projection:
            self.proj_out = nn.Linear(inner_dim, in_channels)
        else:
            self.proj_out = nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0)
        self.adaln_single = None
        self.use_additional_conditions = False
        self.caption_projection = None
        self.gradient_checkpointing = False

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, added_cond_kwargs=None,
                class_labels=None, cross_attention_kwargs=None, attention_mask=None, encoder_attention_mask=None,
                return_dict: bool = True):
        """diffusers 0.23.1 Transformer2DModel.forward, continuous input (restated; the width-only reference class
        inherits it; identical to blocks.py:1221-1308 without the depth gate)."""
        assert attention_mask is None and encoder_attention_mask is None
        batch, _, height, width = hidden_states.shape
        residual = hidden_states
        hidden_states = self.norm(hidden_states)
        if not self.use_linear_projection:
            hidden_states = self.proj_in(hidden_states)
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
        else:
            inner_dim = hidden_states.shape[1]
            hidden_states = hidden_states.permute(0, 2, 3, 1).reshape(batch, height * width, inner_dim)
            hidden_states = self.proj_in(hidden_states)
        for block in self.transformer_blocks:
            hidden_states = block(hidden_states, attention_mask=attention_mask,
                                  encoder_hidden_states=encoder_hidden_states,
                                  encoder_attention_mask=encoder_attention_mask, timestep=timestep,
                                  cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)
        if not self.use_linear_projection:
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
            hidden_states = self.proj_out(hidden_states)
        else:
            hidden_states = self.proj_out(hidden_states)
            hidden_states = hidden_states.reshape(batch, height, width, inner_dim).permute(0, 3, 1, 2).contiguous()
        output = hidden_states + residual
        if not return_dict:
            return (output,)
        return Transformer2DModelOutput(sample=output)


class DualTransformer2DModel(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("dual cross attention is unused by every shipped config")


# ------------------------------------------------------------------------------------------------
# unet_2d_blocks: containers. Constructors keep only what the reference's subclasses do not replace (samplers,
# flags); forwards are the inherited diffusers loops, restated.
# ------------------------------------------------------------------------------------------------
class _DownBase(nn.Module):
    def _common(self, out_channels, add_downsample, downsample_padding):
        self.resnets = nn.ModuleList([])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, use_conv=True, out_channels=out_channels,
                                                        padding=downsample_padding, name="op")]) if add_downsample else None
        self.gradient_checkpointing = False


class CrossAttnDownBlock2D(_DownBase):
    def __init__(self, in_channels, out_channels, temb_channels, dropout=0.0, num_layers=1,
                 transformer_layers_per_block=1, resnet_eps=1e-6, resnet_time_scale_shift="default",
                 resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True, num_attention_heads=1,
                 cross_attention_dim=1280, output_scale_factor=1.0, downsample_padding=1, add_downsample=True,
                 dual_cross_attention=False, use_linear_projection=False, only_cross_attention=False,
                 upcast_attention=False, attention_type="default"):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        self.attentions = nn.ModuleList([])
        self._common(out_channels, add_downsample, downsample_padding)

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None, additional_residuals=None):
        output_states = ()
        blocks = list(zip(self.resnets, self.attentions))
        for i, (resnet, attn) in enumerate(blocks):
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                 encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
            if i == len(blocks) - 1 and additional_residuals is not None:
                hidden_states = hidden_states + additional_residuals
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for downsampler in self.downsamplers:
                hidden_states = downsampler(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class DownBlock2D(_DownBase):
    def __init__(self, in_channels, out_channels, temb_channels, dropout=0.0, num_layers=1, resnet_eps=1e-6,
                 resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True,
                 output_scale_factor=1.0, add_downsample=True, downsample_padding=1):
        super().__init__()
        self._common(out_channels, add_downsample, downsample_padding)

    def forward(self, hidden_states, temb=None, scale: float = 1.0):
        output_states = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb)
            output_states = output_states + (hidden_states,)
        if self.downsamplers is not None:
            for downsampler in self.downsamplers:
                hidden_states = downsampler(hidden_states)
            output_states = output_states + (hidden_states,)
        return hidden_states, output_states


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, in_channels, temb_channels, dropout=0.0, num_layers=1, transformer_layers_per_block=1,
                 resnet_eps=1e-6, resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32,
                 resnet_pre_norm=True, num_attention_heads=1, output_scale_factor=1.0, cross_attention_dim=1280,
                 dual_cross_attention=False, use_linear_projection=False, upcast_attention=False,
                 attention_type="default"):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        self.attentions = nn.ModuleList([])
        self.resnets = nn.ModuleList([])
        self.gradient_checkpointing = False

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, attention_mask=None,
                cross_attention_kwargs=None, encoder_attention_mask=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                 encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
            hidden_states = resnet(hidden_states, temb)
        return hidden_states


class _UpBase(nn.Module):
    def _common(self, out_channels, add_upsample):
        self.resnets = nn.ModuleList([])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, use_conv=True, out_channels=out_channels)]) \
            if add_upsample else None
        self.gradient_checkpointing = False


class CrossAttnUpBlock2D(_UpBase):
    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, dropout=0.0, num_layers=1,
                 transformer_layers_per_block=1, resnet_eps=1e-6, resnet_time_scale_shift="default",
                 resnet_act_fn="swish", resnet_groups=32, resnet_pre_norm=True, num_attention_heads=1,
                 cross_attention_dim=1280, output_scale_factor=1.0, add_upsample=True, dual_cross_attention=False,
                 use_linear_projection=False, only_cross_attention=False, upcast_attention=False,
                 attention_type="default", resolution_idx=None):
        super().__init__()
        self.has_cross_attention = True
        self.num_attention_heads = num_attention_heads
        self.attentions = nn.ModuleList([])
        self.resolution_idx = resolution_idx
        self._common(out_channels, add_upsample)

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                cross_attention_kwargs=None, upsample_size=None, attention_mask=None, encoder_attention_mask=None):
        for resnet, attn in zip(self.resnets, self.attentions):
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
            hidden_states = attn(hidden_states, encoder_hidden_states=encoder_hidden_states,
                                 cross_attention_kwargs=cross_attention_kwargs, attention_mask=attention_mask,
                                 encoder_attention_mask=encoder_attention_mask, return_dict=False)[0]
        if self.upsamplers is not None:
            for upsampler in self.upsamplers:
                hidden_states = upsampler(hidden_states, upsample_size)
        return hidden_states


class UpBlock2D(_UpBase):
    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, dropout=0.0, num_layers=1,
                 resnet_eps=1e-6, resnet_time_scale_shift="default", resnet_act_fn="swish", resnet_groups=32,
                 resnet_pre_norm=True, output_scale_factor=1.0, add_upsample=True, resolution_idx=None):
        super().__init__()
        self.resolution_idx = resolution_idx
        self._common(out_channels, add_upsample)

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, upsample_size=None, scale: float = 1.0):
        for resnet in self.resnets:
            res_hidden_states = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res_hidden_states], dim=1)
            hidden_states = resnet(hidden_states, temb)
        if self.upsamplers is not None:
            for upsampler in self.upsamplers:
                hidden_states = upsampler(hidden_states, upsample_size)
        return hidden_states


@dataclass
class UNet2DConditionOutput:
    sample: torch.FloatTensor = None


# ------------------------------------------------------------------------------------------------
# module assembly
# ------------------------------------------------------------------------------------------------
class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


def _dummy_class(name):
    return type(name, (nn.Module,), {"__init__": lambda self, *a, **k: (_ for _ in ()).throw(
        NotImplementedError(f"diffusers.{name} is not part of the APTP hot path (stub)"))})


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = []
    m.__aptp_shim__ = True
    for k, v in attrs.items():
        setattr(m, k, v)

    def _getattr(attr, _m=m):  # any other imported name: a placeholder class that refuses to be built
        if attr.startswith("__"):
            raise AttributeError(attr)
        c = _dummy_class(attr)
        setattr(_m, attr, c)
        return c
    m.__getattr__ = _getattr
    sys.modules[name] = m
    return m


def install_unet_shim() -> None:
    """Register the stub `diffusers` package tree (replaces the router-only shim if that one is installed)."""
    if getattr(sys.modules.get("diffusers"), "__aptp_unet_shim__", False):
        return
    for k in [k for k in sys.modules if k == "diffusers" or k.startswith("diffusers.")]:
        del sys.modules[k]
    d = _module("diffusers", ModelMixin=ModelMixin, ConfigMixin=ConfigMixin, __version__="0.23.1")
    d.__aptp_unet_shim__ = True
    d.configuration_utils = _module("diffusers.configuration_utils", ConfigMixin=ConfigMixin,
                                    register_to_config=register_to_config)
    d.loaders = _module("diffusers.loaders", UNet2DConditionLoadersMixin=type("UNet2DConditionLoadersMixin", (), {}))
    d.utils = _module("diffusers.utils", logging=_Logging, USE_PEFT_BACKEND=True,
                      _get_model_file=None, _add_variant=None)
    models = _module("diffusers.models", Transformer2DModel=Transformer2DModel,
                     DualTransformer2DModel=DualTransformer2DModel)
    d.models = models
    models.activations = _module("diffusers.models.activations", GEGLU=GEGLU, get_activation=get_activation)
    models.resnet = _module("diffusers.models.resnet", ResnetBlock2D=ResnetBlock2D, Upsample2D=Upsample2D,
                            Downsample2D=Downsample2D)
    models.transformer_2d = _module("diffusers.models.transformer_2d", Transformer2DModelOutput=Transformer2DModelOutput,
                                    Transformer2DModel=Transformer2DModel)
    models.attention = _module("diffusers.models.attention", BasicTransformerBlock=BasicTransformerBlock,
                               FeedForward=FeedForward)
    models.attention_processor = _module(
        "diffusers.models.attention_processor", AttnProcessor2_0=AttnProcessor2_0, Attention=Attention,
        AttnProcessor=AttnProcessor, AttnAddedKVProcessor=AttnAddedKVProcessor, AttentionProcessor=object,
        ADDED_KV_ATTENTION_PROCESSORS=(), CROSS_ATTENTION_PROCESSORS=())
    models.unet_2d_blocks = _module(
        "diffusers.models.unet_2d_blocks", CrossAttnDownBlock2D=CrossAttnDownBlock2D, DownBlock2D=DownBlock2D,
        CrossAttnUpBlock2D=CrossAttnUpBlock2D, UpBlock2D=UpBlock2D, UNetMidBlock2DCrossAttn=UNetMidBlock2DCrossAttn)
    models.embeddings = _module("diffusers.models.embeddings", Timesteps=Timesteps, TimestepEmbedding=TimestepEmbedding)
    # pdm/utils/op_counter.py keys its hook table on these classes; with the stubs every conv / linear is a plain
    # nn.Conv2d / nn.Linear (same hook formulas, op_counter.py:60-116), so placeholders suffice
    models.normalization = _module("diffusers.models.normalization")
    models.lora = _module("diffusers.models.lora")
    models.unet_2d_condition = _module("diffusers.models.unet_2d_condition", UNet2DConditionOutput=UNet2DConditionOutput)
    # `from diffusers.models.modeling_utils import *` is how unet_2d_conditional.py gets os / typing names
    models.modeling_utils = _module(
        "diffusers.models.modeling_utils", ModelMixin=ModelMixin, _LOW_CPU_MEM_USAGE_DEFAULT=False, os=os, torch=torch,
        nn=nn, Any=Any, Callable=Callable, Dict=Dict, List=List, Optional=Optional, Tuple=Tuple, Union=Union)
    models.modeling_utils.__all__ = ["ModelMixin", "os", "torch", "nn", "Any", "Callable", "Dict", "List", "Optional",
                                     "Tuple", "Union"]


def load_unet_reference(root: str = "/root/reference") -> Dict[str, types.ModuleType]:
    """Execute the reference's own gates.py / estimation_utils.py / hypernet.py / blocks.py / unet_2d_conditional.py in
    place (never copied) on top of the stubs. Returns {"blocks": ..., "unet": ..., "gates": ..., "hypernet": ...}."""
    import importlib.util
    install_unet_shim()
    for k in [k for k in sys.modules if k == "pdm" or k.startswith("pdm.")]:
        del sys.modules[k]
    for pkg in ("pdm", "pdm.utils", "pdm.models", "pdm.models.unet", "pdm.models.hypernet", "pdm.losses"):
        m = types.ModuleType(pkg)
        m.__path__ = []
        sys.modules[pkg] = m

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, f"{root}/{rel}")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    est = load("pdm.utils.estimation_utils", "pdm/utils/estimation_utils.py")
    gates = load("pdm.models.unet.gates", "pdm/models/unet/gates.py")
    hyp = load("pdm.models.hypernet.hypernet", "pdm/models/hypernet/hypernet.py")
    sys.modules["pdm.models.hypernet"].HyperStructure = hyp.HyperStructure
    blocks = load("pdm.models.unet.blocks", "pdm/models/unet/blocks.py")
    out = {"estimation_utils": est, "gates": gates, "hypernet": hyp, "blocks": blocks, "unet": None, "unet_error": None,
           "op_counter": None}
    try:
        out["op_counter"] = load("pdm.utils.op_counter", "pdm/utils/op_counter.py")
    except Exception as e:  # noqa: BLE001
        out["op_counter_error"] = repr(e)
    try:
        out["unet"] = load("pdm.models.unet.unet_2d_conditional", "pdm/models/unet/unet_2d_conditional.py")
    except Exception as e:  # noqa: BLE001 -- the layer goldens do not depend on the top-level file
        out["unet_error"] = repr(e)
    return out
