"""Closed form of the reference's MAC accounting (it enters the resource loss).

The reference stamps `__macs__` on every leaf with forward hooks during one batch-1 forward
(pdm/utils/op_counter.py:19-116, :259-306) and then walks the module tree
(pdm/models/unet/unet_2d_conditional.py:2124-2163 and the per-block calc_macs methods). Both steps are
fixed arithmetic in the layer shapes and the hard gate bits, evaluated here directly (SURVEY
Appendix F). Quirks kept on purpose because they enter the loss: Linear bias counted once per call,
norms/SiLU modules counted as MACs, cross-attention counted with the *query* length for both matmuls.
Ratios carry the straight-through gradient of hard_concrete (pdm/utils/estimation_utils.py:67-75).
"""
from __future__ import annotations

from typing import Any, Dict, List

import torch


def hard_concrete(x: torch.Tensor) -> torch.Tensor:
    hard = (x >= 0.5).to(x.dtype)
    return (hard - x).detach() + x


def _lin(n_tokens: int, cin: int, cout: int, bias: bool) -> float:
    return float(n_tokens * cin * cout + (cout if bias else 0))  # op_counter.py:60-65


def _conv(k: int, cin: int, cout: int, hw_out: int) -> float:
    return float(k * k * cin * cout * hw_out + cout * hw_out)  # op_counter.py:89-116 (bias present)


def _resnet_macs(r, hw: int):
    tdim = r.time_emb_proj.in_features
    prunable = _conv(3, r.cin, r.cout, hw) + _lin(1, tdim, r.cout, True) + 2.0 * r.cout * hw + _conv(3, r.cout, r.cout, hw)
    total = prunable + 2.0 * r.cin * hw  # + norm1 (blocks.py:384-409)
    if r.conv_shortcut is not None:
        total += _conv(1, r.cin, r.cout, hw)
    return prunable, total


def _attn_macs(dim: int, heads: int, n_q: int, n_ctx: int, ctx_dim: int) -> float:
    hd = dim // heads
    m = _lin(n_q, dim, dim, False) + 2 * _lin(n_ctx, ctx_dim, dim, False)
    m += heads * (2.0 * n_q * n_q * hd + n_q * n_q)  # op_counter.py:286-297 (seq_len = output length)
    m += _lin(n_q, dim, dim, True)
    return m


def _transformer_macs(t, hw: int, n_ctx: int):
    d = t.dim
    a1 = _attn_macs(d, t.heads, hw, hw, d)
    a2 = _attn_macs(d, t.heads, hw, n_ctx, t.ctx_dim)
    ff = _lin(hw, d, 8 * d, True) + _lin(hw, 4 * d, d, True)  # blocks.py:103-112
    fixed = 2.0 * d * hw + 2 * _lin(hw, d, d, True) + 3.0 * hw * d  # GroupNorm + proj_in/out + 3 LayerNorm
    return {"attn1": a1, "attn2": a2, "ff": ff, "fixed": fixed}


def build_resource_info(model, H: int, W: int, n_ctx: int = 77) -> Dict[str, Any]:
    from .unet import ResnetBlock2DWidthGated
    cfg = model.config
    ch = cfg["block_out_channels"]
    info: List[tuple] = []
    fixed = 0.0
    te = model.time_embedding
    fixed += _lin(1, te.linear_1.in_features, te.linear_1.out_features, True)
    fixed += 2.0 * te.linear_1.out_features  # nn.SiLU hook: 2 * numel (op_counter.py:55-57)
    fixed += _lin(1, te.linear_2.in_features, te.linear_2.out_features, True)
    h, w = H, W
    fixed += _conv(3, cfg["in_channels"], ch[0], h * w)

    def add_block(blk, hw):
        for r in blk.resnets:
            info.append(("res", r, _resnet_macs(r, hw)))
        if blk.attentions is not None:
            for a in blk.attentions:
                info.append(("attn", a, _transformer_macs(a, hw, n_ctx)))

    for blk in model.down_blocks:
        add_block(blk, h * w)
        if blk.downsamplers is not None:
            h, w = h // 2, w // 2
            c = blk.downsamplers[0].conv.in_channels
            fixed += _conv(3, c, c, h * w)
    mb = model.mid_block  # get_structure order for the mid block is resnets then attentions too
    info.append(("res", mb.resnets[0], _resnet_macs(mb.resnets[0], h * w)))
    info.append(("res", mb.resnets[1], _resnet_macs(mb.resnets[1], h * w)))
    info.append(("attn", mb.attentions[0], _transformer_macs(mb.attentions[0], h * w, n_ctx)))
    for blk in model.up_blocks:
        add_block(blk, h * w)
        if blk.upsamplers is not None:
            h, w = h * 2, w * 2
            c = blk.upsamplers[0].conv.in_channels
            fixed += _conv(3, c, c, h * w)
    fixed += 2.0 * ch[0] * h * w + 2.0 * ch[0] * h * w  # conv_norm_out + conv_act
    fixed += _conv(3, ch[0], cfg["out_channels"], h * w)
    return {"layers": info, "fixed": fixed, "H": H, "W": W}


def _ratio(g: torch.Tensor) -> torch.Tensor:
    hg = hard_concrete(g)
    return hg.sum(dim=1, keepdim=True) / hg.shape[1]


def calc_macs(model) -> Dict[str, Any]:
    ri = model._macs_table
    total, prunable = ri["fixed"], 0.0
    cur_p, cur_t = 0.0, ri["fixed"]
    for kind, m, mm in ri["layers"]:
        if kind == "res":
            P, T = mm
            r = _ratio(m.gate.gate_f)
            cp = r * P
            ct = r.detach() * P + (T - P)
            if m.depth_gate is not None:  # blocks.py:626-633
                d = hard_concrete(m.depth_gate.gate_f).unsqueeze(1)
                cp = (r * P + (T - P)) * d
                ct = ct * d.detach()
        else:
            tb = m.transformer_blocks[0]
            P = mm["attn1"] + mm["attn2"] + mm["ff"]
            T = P + mm["fixed"]
            r1, r2, rf = _ratio(tb.attn1.gate.gate_f), _ratio(tb.attn2.gate.gate_f), _ratio(tb.ff.net[0].gate.gate_f)
            cp = r1 * mm["attn1"] + r2 * mm["attn2"] + rf * mm["ff"]
            ct = r1.detach() * mm["attn1"] + r2.detach() * mm["attn2"] + rf.detach() * mm["ff"] + mm["fixed"]
            if m.depth_gate is not None:  # blocks.py:1400-1411
                d = hard_concrete(m.depth_gate.gate_f).unsqueeze(1)
                cp = (cp + T - P) * d
                ct = ct * d.detach()
        total += T
        prunable += P
        cur_p = cur_p + cp
        cur_t = cur_t + ct
    return {"total_macs": total, "prunable_macs": prunable, "cur_prunable_macs": cur_p, "cur_total_macs": cur_t}


def prunable_macs_list(model) -> List[List[float]]:
    """get_prunable_macs (unet_2d_conditional.py:2165-2172): one list per gated sub-block, in
    get_structure order (consumed by StructureVectorQuantizer.set_prunable_macs_template)."""
    out = []
    for kind, m, mm in model._macs_table["layers"]:
        if kind == "res":
            out.append([mm[0]])
        else:
            out.append([mm["attn1"], mm["attn2"], mm["ff"]])
    return out


def block_utilization(model):
    """get_block_utilization (unet_2d_conditional.py:2174-2181) flattened per gated sub-block."""
    util = []
    for kind, m, mm in model._macs_table["layers"]:
        if kind == "res":
            u = hard_concrete(m.gate.gate_f).mean(dim=1)
        else:
            tb = m.transformer_blocks[0]
            tot = mm["attn1"] + mm["attn2"] + mm["ff"]
            u = (hard_concrete(tb.attn1.gate.gate_f).mean(1) * mm["attn1"] +
                 hard_concrete(tb.attn2.gate.gate_f).mean(1) * mm["attn2"] +
                 hard_concrete(tb.ff.net[0].gate.gate_f).mean(1) * mm["ff"]) / tot
        if m.depth_gate is not None:
            u = u * hard_concrete(m.depth_gate.gate_f)
        util.append(u)
    return util
