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


def _calc_macs_torch(model) -> Dict[str, Any]:
    """The closed form in torch ops, sub-block by sub-block (host-side accounting on CPU gate tensors:
    count_macs() at construction time, CPU tests of the closed form against the reference's hook formulas)."""
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



def _device_tables(model, device):
    """aptp_macs_gate / aptp_macs_sub records (include/aptp_sm100.h) of the current resource table, resident on
    `device`; sub-blocks in get_structure order, width columns first then the depth columns."""
    import numpy as np

    from . import kernels as K
    ri = model._macs_table
    cache = getattr(model, "_macs_dev", None)
    if cache is not None and cache["table"] is ri and cache["device"] == device:
        return cache
    by_id = {id(m): (kind, mm) for kind, m, mm in ri["layers"]}
    n_width = sum(w for m in model._gated for w in m.gate_widths())
    gates, subs = [], []
    col = di = 0
    total, prunable = ri["fixed"], 0.0
    for m in model._gated:
        kind, mm = by_id[id(m)]
        if kind == "res":
            macs, fixed = [mm[0]], mm[1] - mm[0]
        else:
            macs, fixed = [mm["attn1"], mm["attn2"], mm["ff"]], mm["fixed"]
        first = len(gates)
        for wd, mc in zip(m.gate_widths(), macs):
            gates.append((col, wd, mc))
            col += wd
        dcol = -1
        if m.depth_gate is not None:
            dcol = n_width + di
            di += 1
        subs.append((first, len(macs), dcol, 0, fixed))
        prunable += sum(macs)
        total += sum(macs) + fixed
    g_np = np.array(gates, dtype=K.MACS_GATE_DTYPE)
    s_np = np.array(subs, dtype=K.MACS_SUB_DTYPE)
    cache = {"table": ri, "device": device, "n_gates": len(gates), "n_subs": len(subs), "dim": n_width + di,
             "gates": torch.from_numpy(g_np.view(np.uint8).copy()).to(device),
             "subs": torch.from_numpy(s_np.view(np.uint8).copy()).to(device),
             "fixed": ri["fixed"], "total": total, "prunable": prunable}
    model._macs_dev = cache
    return cache


class _MacsRatio(torch.autograd.Function):
    """(cur_prunable_macs, cur_total_macs) [B, 1] from the [B, 1620] gate matrix: aptp_macs_ratio_fwd / _bwd."""

    @staticmethod
    def forward(ctx, arch: torch.Tensor, tab):
        from . import kernels as K
        arch = arch.detach().contiguous()
        B = arch.shape[0]
        cur_p = torch.empty(B, device=arch.device, dtype=torch.float32)
        cur_t = torch.empty(B, device=arch.device, dtype=torch.float32)
        K.macs_ratio_fwd(arch, tab["gates"], tab["n_gates"], tab["subs"], tab["n_subs"], tab["fixed"], cur_p, cur_t)
        ctx.save_for_backward(arch)
        ctx.tab = tab
        cur_p, cur_t = cur_p.unsqueeze(1), cur_t.unsqueeze(1)
        ctx.mark_non_differentiable(cur_t)
        return cur_p, cur_t

    @staticmethod
    def backward(ctx, dcur_p, _dcur_t):
        from . import kernels as K
        (arch,) = ctx.saved_tensors
        tab = ctx.tab
        darch = torch.empty_like(arch)
        K.macs_ratio_bwd(arch, tab["gates"], tab["n_gates"], tab["subs"], tab["n_subs"],
                         dcur_p.reshape(-1).to(torch.float32).contiguous(), darch)
        return darch, None


def calc_macs(model) -> Dict[str, Any]:
    """unet_2d_conditional.py:2124-2163. Gates resident on the GPU (every call of the train / sampling path): one
    kernel over the concatenated gate matrix; CPU gate tensors: the same closed form in torch ops."""
    flat_w, flat_d = model._flat_gates
    if not flat_w[0].is_cuda:
        return _calc_macs_torch(model)
    tab = _device_tables(model, flat_w[0].device)
    rows = flat_w[0].shape[0]
    arch = torch.cat([w.reshape(rows, -1).to(torch.float32) for w in flat_w] +
                     [d.reshape(rows, 1).to(torch.float32) for d in flat_d], dim=1)
    assert arch.shape[1] == tab["dim"], (arch.shape, tab["dim"])
    cur_p, cur_t = _MacsRatio.apply(arch, tab)
    return {"total_macs": tab["total"], "prunable_macs": tab["prunable"], "cur_prunable_macs": cur_p,
            "cur_total_macs": cur_t}


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
