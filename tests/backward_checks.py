"""Backward-kernel numerics (K5): each kernel against torch autograd of the plain fp32 op on the same
(bf16-rounded) inputs. Used by tests/test_backward_gpu.py and tools/kernel_check.py."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from diffusion_pruning_b200 import kernels as K
from kernel_checks import DEV, _close, _rand


def check_scale_cols(B=3, hw=200, C=384, group=64, seed=0):
    u = _rand(B * hw, C, seed=seed).bfloat16()
    dy = _rand(B * hw, C, seed=seed + 1).bfloat16()
    gate = torch.rand(B, C // group, device=DEV) * 0.9 + 0.05
    y = torch.empty_like(u)
    K.scale_cols(u, C, y, C, B, hw, C, gate, C // group, group)
    gfull = gate.repeat_interleave(hw, 0).repeat_interleave(group, 1)
    _close(y, u.float() * gfull, 1e-2, 1e-2, "scale_cols fwd")
    du = torch.empty_like(u)
    dg = torch.zeros(B, C // group, device=DEV)
    K.scale_cols_bwd(u, C, dy, C, du, C, B, hw, C, gate, C // group, group, dg)
    _close(du, dy.float() * gfull, 1e-2, 1e-2, "scale_cols du")
    ref = (dy.float() * u.float()).reshape(B, hw, C // group, group).sum(dim=(1, 3))
    _close(dg, ref, 5e-2, 2e-3, "scale_cols dgate")


def check_geglu_train(B=2, hw=130, inner=2560, gw=32, use_gate=True, seed=0):
    hg = _rand(B * hw, 2 * inner, seed=seed).bfloat16()
    df = _rand(B * hw, inner, seed=seed + 1).bfloat16()
    gate = (torch.rand(B, gw, device=DEV) * 0.9 + 0.05) if use_gate else None
    grp = inner // gw
    out = torch.empty(B * hw, inner, device=DEV, dtype=torch.bfloat16)
    K.geglu(hg, 2 * inner, out, inner, B, hw, inner, gate, gw, grp)
    x = hg.float().requires_grad_(True)
    g = gate.clone().requires_grad_(True) if use_gate else None
    gfull = g.repeat_interleave(hw, 0).repeat_interleave(grp, 1) if use_gate else 1.0
    ref = (gfull * x[:, :inner]) * F.gelu(gfull * x[:, inner:])
    _close(out, ref.detach(), 2e-2, 1e-2, "geglu_train fwd")
    ref.backward(df.float())
    dhg = torch.empty_like(hg)
    dg = torch.zeros(B, gw, device=DEV)
    K.geglu_bwd(hg, 2 * inner, df, inner, dhg, 2 * inner, B, hw, inner, gate, gw, grp, dg if use_gate else None)
    _close(dhg, x.grad, 2e-2, 1e-2, "geglu_train dhg")
    if use_gate:
        _close(dg, g.grad, 0.5, 5e-3, "geglu_train dgate")


def check_groupnorm_bwd(B=3, hw=256, C=320, groups=32, silu=True, use_gate=True, accumulate=False, seed=0):
    gs = C // groups
    x = (_rand(B * hw, C, seed=seed) * 2 + 0.5).bfloat16()
    da = _rand(B * hw, C, seed=seed + 1).bfloat16()
    gamma = _rand(C, seed=seed + 2) * 0.2 + 1
    beta = _rand(C, seed=seed + 3) * 0.2
    gate = (torch.rand(B, groups, device=DEV) * 0.9 + 0.05) if use_gate else None
    stats = torch.zeros(B, groups, 2, device=DEV)
    K.groupnorm_stats(x, C, C, None, 0, 0, B, hw, gs, None, stats, groups)
    dx0 = _rand(B * hw, C, seed=seed + 4).bfloat16()
    dx = dx0.clone() if accumulate else torch.full_like(x, float("nan"))
    bstats = torch.empty(B, groups, 2, device=DEV)
    dg = torch.zeros(B, groups, device=DEV)
    K.groupnorm_bwd(x, C, da, C, dx, C, accumulate, B, hw, C, gs, 1e-5, stats, groups, gamma, beta, gate, groups, silu,
                    bstats, dg if use_gate else None)
    xr = x.float().requires_grad_(True)
    g = gate.clone().requires_grad_(True) if use_gate else None
    xin = xr.reshape(B, hw, C).permute(0, 2, 1)
    if use_gate:
        xin = xin * g.repeat_interleave(gs, 1)[:, :, None]
    y = F.group_norm(xin, groups, gamma, beta, 1e-5)
    if silu:
        y = F.silu(y)
    y.backward(da.float().reshape(B, hw, C).permute(0, 2, 1))
    ref = xr.grad + (dx0.float() if accumulate else 0)
    _close(dx, ref, 3e-2, 2e-2, f"groupnorm_bwd dx silu={silu} gate={use_gate}")
    if use_gate:
        _close(dg, g.grad, 0.3, 2e-2, "groupnorm_bwd dgate")


def check_groupnorm_affine(B=3, hw=200, C=320, groups=32, silu=True, use_gate=False, seed=5):
    """aptp_groupnorm_bwd_affine: dx as before plus (dgamma, dbeta) vs autograd."""
    gs = C // groups
    x = (_rand(B * hw, C, seed=seed) * 2 + 0.5).bfloat16()
    da = _rand(B * hw, C, seed=seed + 1).bfloat16()
    gamma = _rand(C, seed=seed + 2) * 0.2 + 1
    beta = _rand(C, seed=seed + 3) * 0.2
    gate = (torch.rand(B, groups, device=DEV) * 0.9 + 0.05) if use_gate else None
    stats = torch.zeros(B, groups, 2, device=DEV)
    K.groupnorm_stats(x, C, C, None, 0, 0, B, hw, gs, None, stats, groups)
    dx = torch.full_like(x, float("nan"))
    bstats = torch.empty(B, groups, 2, device=DEV)
    dg = torch.zeros(B, groups, device=DEV)
    daff = torch.zeros(C, 2, device=DEV)
    K.groupnorm_bwd_affine(x, C, da, C, dx, C, False, B, hw, C, gs, 1e-5, stats, groups, gamma, beta, gate, groups, silu,
                           bstats, dg if use_gate else None, daff)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    xin = xr.reshape(B, hw, C).permute(0, 2, 1)
    if use_gate:
        xin = xin * gate.repeat_interleave(gs, 1)[:, :, None]
    y = F.group_norm(xin, groups, gr, br, 1e-5)
    if silu:
        y = F.silu(y)
    y.backward(da.float().reshape(B, hw, C).permute(0, 2, 1))
    _close(dx, xr.grad, 3e-2, 2e-2, "groupnorm_bwd_affine dx")
    _close(daff[:, 0], gr.grad, 2e-2 * gr.grad.abs().max().item(), 1e-2, "groupnorm dgamma")
    _close(daff[:, 1], br.grad, 2e-2 * br.grad.abs().max().item(), 1e-2, "groupnorm dbeta")


def check_layernorm_affine(rows=777, C=640, seed=6):
    x = (_rand(rows, C, seed=seed) * 3 + 1).bfloat16()
    dy = _rand(rows, C, seed=seed + 1).bfloat16()
    daff = torch.zeros(C, 2, device=DEV)
    K.layernorm_affine_bwd(x, C, dy, C, rows, C, 1e-5, daff)
    xr = x.float()
    g = torch.ones(C, device=DEV, requires_grad=True)
    b = torch.zeros(C, device=DEV, requires_grad=True)
    F.layer_norm(xr, (C,), g, b, 1e-5).backward(dy.float())
    _close(daff[:, 0], g.grad, 2e-3 * g.grad.abs().max().item(), 2e-3, "layernorm dgamma")
    _close(daff[:, 1], b.grad, 2e-3 * b.grad.abs().max().item(), 2e-3, "layernorm dbeta")


def check_layernorm_bwd(rows=777, C=640, accumulate=True, seed=0):
    x = (_rand(rows, C, seed=seed) * 3 + 1).bfloat16()
    dy = _rand(rows, C, seed=seed + 1).bfloat16()
    gamma = _rand(C, seed=seed + 2) * 0.2 + 1
    dx0 = _rand(rows, C, seed=seed + 3).bfloat16()
    dx = dx0.clone() if accumulate else torch.full_like(x, float("nan"))
    K.layernorm_bwd(x, C, dy, C, dx, C, accumulate, rows, C, 1e-5, gamma)
    xr = x.float().requires_grad_(True)
    F.layer_norm(xr, (C,), gamma, torch.zeros_like(gamma), 1e-5).backward(dy.float())
    _close(dx, xr.grad + (dx0.float() if accumulate else 0), 3e-2, 2e-2, "layernorm_bwd")


def check_depth_lerp_bwd(B=3, hw=100, C=128, seed=0):
    x = _rand(B * hw, C + 64, seed=seed).bfloat16()
    y = _rand(B * hw, C, seed=seed + 1).bfloat16()
    dout = _rand(B * hw, C, seed=seed + 2).bfloat16()
    d = torch.tensor([0.25, 1.0, 0.6], device=DEV)[:B]
    dy = torch.empty_like(y)
    dx0 = _rand(B * hw, C + 64, seed=seed + 3).bfloat16()
    dx = dx0.clone()
    dd = torch.zeros(B, device=DEV)
    K.depth_lerp_bwd(dout, C, x, C + 64, y, C, dy, C, dx, C + 64, True, B, hw, C, d, dd)
    df = d.repeat_interleave(hw)[:, None]
    _close(dy, df * dout.float(), 1e-2, 1e-2, "depth_lerp_bwd dy")
    _close(dx[:, :C], dx0[:, :C].float() + (1 - df) * dout.float(), 2e-2, 1e-2, "depth_lerp_bwd dx")
    assert torch.equal(dx[:, C:], dx0[:, C:]), "depth_lerp_bwd touched columns beyond C"
    ref = (dout.float() * (y.float() - x[:, :C].float())).reshape(B, -1).sum(1)
    _close(dd, ref, 0.5, 5e-3, "depth_lerp_bwd dd")


def check_resample_bwd(B=2, H=8, W=8, C=64, seed=0):
    dy = _rand(B * 4 * H * W, C, seed=seed).bfloat16()
    dx = torch.empty(B * H * W, C, device=DEV, dtype=torch.bfloat16)
    K.upsample2x_bwd(dy, dx, B, H, W, C)
    ref = dy.float().reshape(B, H, 2, W, 2, C).sum(dim=(2, 4)).reshape(B * H * W, C)
    _close(dx, ref, 2e-2, 1e-2, "upsample2x_bwd")
    src = _rand(B * H * W, C, seed=seed + 1).bfloat16()
    dst = torch.full((B * 4 * H * W, C), 7.0, device=DEV, dtype=torch.bfloat16)
    K.zero_insert2x(src, dst, B, H, W, C)
    r = torch.zeros(B, 2 * H, 2 * W, C, device=DEV)
    r[:, ::2, ::2] = src.float().reshape(B, H, W, C)
    _close(dst, r.reshape(-1, C), 0, 0, "zero_insert2x")
    a = _rand(B * H * W, C, seed=seed + 2).bfloat16()
    b = _rand(B * H * W, C + 8, seed=seed + 3).bfloat16()
    b0 = b.clone()
    K.add_rows(a, C, b, C + 8, B * H * W, C)
    _close(b[:, :C], a.float() + b0[:, :C].float(), 2e-2, 1e-2, "add_rows")
    assert torch.equal(b[:, C:], b0[:, C:])


def check_conv_dgrad(B=2, H=16, W=16, Cin=128, Cout=192, stride=1, seed=0):
    """dgrad of the 3x3 conv through the SAME grouped GEMM kernel: transposed, tap-flipped packed weights
    (stride 2: zero-inserted output gradient first)."""
    from diffusion_pruning_b200._lib import A_CONV3X3
    from diffusion_pruning_b200.plan import pack_conv_weight_dgrad
    w = _rand(Cout, Cin, 3, 3, scale=(9 * Cin) ** -0.5, seed=seed).bfloat16()
    Ho, Wo = H // stride, W // stride
    dy = _rand(B, Cout, Ho, Wo, seed=seed + 1).bfloat16()
    x = torch.zeros(B, Cin, H, W, device=DEV, requires_grad=True)
    F.conv2d(x, w.float(), None, stride=stride, padding=1).backward(dy.float())
    wt = pack_conv_weight_dgrad(w)  # [Cin, 9*Cout]
    dy_rows = dy.permute(0, 2, 3, 1).reshape(B * Ho * Wo, Cout).contiguous()
    if stride == 2:
        up = torch.empty(B * H * W, Cout, device=DEV, dtype=torch.bfloat16)
        K.zero_insert2x(dy_rows, up, B, Ho, Wo, Cout)
        dy_rows = up
    dx = torch.full((B * H * W, Cin), float("nan"), device=DEV, dtype=torch.bfloat16)
    sched = K.build_schedule([K.Segment(0, B * H * W, Cin, (Cout + 63) // 64)], 128, DEV, mode=A_CONV3X3, Ho=H, Wo=W)
    K.grouped_gemm(dy_rows, wt, dx, sched, a_ld=Cout, a_k=Cout, a_rows=B * H * W, mode=A_CONV3X3, batch=B, H=H, W=W,
                   k_tap_pitch=Cout, out_ld=Cin, rows_per_sample=H * W)
    K.check_abort()
    _close(dx, x.grad.permute(0, 2, 3, 1).reshape(B * H * W, Cin), 3e-2, 2e-2, f"conv dgrad stride {stride}")


def check_attention_bwd(B=2, heads=3, kept=(3, 1), Nq=256, Nkv=256, seed=0):
    C = heads * 64
    q = _rand(B * Nq, C, seed=seed).bfloat16()
    k = _rand(B * Nkv, C, seed=seed + 1).bfloat16()
    v = _rand(B * Nkv, C, seed=seed + 2).bfloat16()
    do = _rand(B * Nq, C, seed=seed + 3).bfloat16()
    out = torch.zeros(B * Nq, C, device=DEV, dtype=torch.bfloat16)
    sh = torch.tensor(list(kept), device=DEV, dtype=torch.int32)
    lse = torch.zeros(B, heads, Nq, device=DEV)
    K.attention(q, C, k, C, v, C, out, C, B, Nq, Nkv, sh, heads, 0.125, lse)
    delta = torch.empty(B, heads, Nq, device=DEV)
    dq = torch.full_like(q, 9.0)
    dk = torch.full_like(k, 9.0)
    dv = torch.full_like(v, 9.0)
    K.attention_bwd(q, C, k, C, v, C, out, C, do, C, lse, delta, dq, C, dk, C, dv, C, B, Nq, Nkv, sh, heads, 0.125)
    K.check_abort()
    qf = q.float().reshape(B, Nq, heads, 64).transpose(1, 2).requires_grad_(True)
    kf = k.float().reshape(B, Nkv, heads, 64).transpose(1, 2).requires_grad_(True)
    vf = v.float().reshape(B, Nkv, heads, 64).transpose(1, 2).requires_grad_(True)
    sc = (qf @ kf.transpose(-1, -2)) * 0.125
    ref = torch.softmax(sc, -1) @ vf
    ref.backward(do.float().reshape(B, Nq, heads, 64).transpose(1, 2))
    lse_ref = torch.logsumexp(sc.detach(), -1) * 1.4426950408889634
    for b in range(B):
        kb = kept[b]
        _close(lse[b, :kb], lse_ref[b, :kb], 2e-2, 1e-3, f"attention lse sample {b}")
        for name, got, r, n in (("dq", dq, qf.grad, Nq), ("dk", dk, kf.grad, Nkv), ("dv", dv, vf.grad, Nkv)):
            g = got.reshape(B, n, heads, 64).float()[b]
            rr = r[b].transpose(0, 1)
            scale = rr[:, :kb].abs().max().item() if kb else 1.0
            _close(g[:, :kb], rr[:, :kb], 3e-2 * max(scale, 1e-3), 3e-2, f"attention_bwd {name} sample {b} Nq{Nq} Nkv{Nkv}")
            assert (g[:, kb:] == 9.0).all(), f"attention_bwd {name}: pruned heads must not be written"


def check_wgrad_linear(rows=1000, n_out=200, k_in=136, ld_extra=8, splits=0, seed=0):
    """aptp_wgrad (linear): dW = dY^T A and db = column sums vs fp32 torch on the same bf16 operands; ragged rows /
    columns (not multiples of 128 / 64), pitched inputs, split-K through fp32 atomics."""
    ldy, lda = n_out + ld_extra, k_in + ld_extra
    dy = _rand(rows, ldy, seed=seed).bfloat16()
    a = _rand(rows, lda, seed=seed + 1).bfloat16()
    dw = torch.zeros(n_out, k_in, device=DEV)
    db = torch.zeros(n_out, device=DEV)
    K.wgrad(dy, ldy, a, lda, dw, db, rows, n_out, k_in, splits=splits)
    K.check_abort()
    ref = dy[:, :n_out].float().t() @ a[:, :k_in].float()
    scale = ref.abs().max().item()
    _close(dw, ref, 2e-3 * scale, 2e-3, f"wgrad linear rows{rows} n{n_out} k{k_in} splits{splits}")
    _close(db, dy[:, :n_out].float().sum(0), 2e-3 * rows ** 0.5, 2e-3, "wgrad bias")


def check_wgrad_conv(B=2, H=16, W=16, cin=72, cout=136, splits=0, seed=3):
    """aptp_wgrad (3x3 conv, stride 1, zero padding): vs autograd of F.conv2d w.r.t. the weight, OHWI layout."""
    rows = B * H * W
    a = _rand(rows, cin, seed=seed).bfloat16()
    dy = _rand(rows, cout, seed=seed + 1).bfloat16()
    dw = torch.zeros(cout, 9 * cin, device=DEV)
    db = torch.zeros(cout, device=DEV)
    K.wgrad(dy, cout, a, cin, dw, db, rows, cout, cin, conv=(B, H, W), splits=splits)
    K.check_abort()
    x = a.float().reshape(B, H, W, cin).permute(0, 3, 1, 2)
    w = torch.zeros(cout, cin, 3, 3, device=DEV, requires_grad=True)
    y = torch.nn.functional.conv2d(x, w, padding=1)
    y.backward(dy.float().reshape(B, H, W, cout).permute(0, 3, 1, 2))
    ref = w.grad.permute(0, 2, 3, 1).reshape(cout, 9 * cin)  # OIHW -> OHWI
    scale = ref.abs().max().item()
    _close(dw, ref, 2e-3 * scale, 2e-3, f"wgrad conv B{B} {H}x{W} cin{cin} cout{cout} splits{splits}")
    _close(db, dy.float().sum(0), 2e-3 * rows ** 0.5, 2e-3, "wgrad conv bias")


def check_wgrad_conv_s2(B=2, H=16, W=16, C=128, cout=72, seed=7):
    """aptp_wgrad, 3x3 stride-2 conv (the down-samplers) vs autograd of F.conv2d(stride=2, padding=1)."""
    Ho, Wo = H // 2, W // 2
    a = _rand(B * H * W, C, seed=seed).bfloat16()
    dy = _rand(B * Ho * Wo, cout + 8, seed=seed + 1).bfloat16()
    dw = torch.zeros(cout, 9 * C, device=DEV)
    K.wgrad(dy, cout + 8, a, C, dw, None, B * Ho * Wo, cout, C, conv=(B, H, W), stride=2)
    K.check_abort()
    x = a.float().reshape(B, H, W, C).permute(0, 3, 1, 2)
    w = torch.zeros(cout, C, 3, 3, device=DEV, requires_grad=True)
    y = torch.nn.functional.conv2d(x, w, stride=2, padding=1)
    y.backward(dy[:, :cout].float().reshape(B, Ho, Wo, cout).permute(0, 3, 1, 2))
    ref = w.grad.permute(0, 2, 3, 1).reshape(cout, 9 * C)
    _close(dw, ref, 2e-3 * ref.abs().max().item(), 2e-3, "wgrad conv stride 2")


def check_col_sum_groups(B=3, hw=5000, C=136, seed=8):
    dy = _rand(B * hw, C + 8, seed=seed).bfloat16()
    out = torch.zeros(B, C, device=DEV)
    K.col_sum_groups(dy, C + 8, B, hw, C, out)
    ref = dy[:, :C].float().reshape(B, hw, C).sum(1)
    _close(out, ref, 2e-3 * hw ** 0.5, 2e-3, "col_sum_groups")


ALL = [
    ("wgrad_conv_stride2", check_wgrad_conv_s2),
    ("wgrad_conv_stride2_64", lambda: check_wgrad_conv_s2(B=1, H=64, W=64, C=64, cout=64)),
    ("col_sum_groups", check_col_sum_groups),
    ("affine_groupnorm", check_groupnorm_affine),
    ("affine_groupnorm_gate_plain", lambda: check_groupnorm_affine(B=2, hw=64, C=1280, silu=False, use_gate=True)),
    ("affine_layernorm", check_layernorm_affine),
    ("affine_layernorm_320", lambda: check_layernorm_affine(rows=4100, C=320)),
    ("affine_layernorm_1280", lambda: check_layernorm_affine(rows=64, C=1280)),
    ("wgrad_linear", check_wgrad_linear),
    ("wgrad_linear_1split", lambda: check_wgrad_linear(rows=4096, n_out=320, k_in=320, ld_extra=0, splits=1)),
    ("wgrad_linear_small", lambda: check_wgrad_linear(rows=77, n_out=64, k_in=1024, splits=0)),
    ("wgrad_conv", check_wgrad_conv),
    ("wgrad_conv_8x8", lambda: check_wgrad_conv(B=3, H=8, W=8, cin=256, cout=128, splits=2)),
    ("wgrad_conv_64", lambda: check_wgrad_conv(B=1, H=64, W=64, cin=64, cout=64, splits=4)),
    ("bwd_attention", check_attention_bwd),
    ("bwd_attention_cross77", lambda: check_attention_bwd(Nkv=77)),
    ("bwd_attention_small", lambda: check_attention_bwd(B=3, heads=2, kept=(2, 0, 1), Nq=64, Nkv=64)),
    ("bwd_attention_long", lambda: check_attention_bwd(B=1, heads=1, kept=(1,), Nq=1024, Nkv=640)),
    # ragged tiles: query count not a multiple of the 64-query tile (per-warp LSE / delta slots fold the bounds in),
    # key count not a multiple of 128 / 256 (one- and two-tile CTAs with a partial tile)
    ("bwd_attention_ragged", lambda: check_attention_bwd(B=2, heads=2, kept=(2, 1), Nq=200, Nkv=328)),
    ("bwd_attention_576", lambda: check_attention_bwd(B=1, heads=2, kept=(2,), Nq=576, Nkv=576)),
    ("bwd_attention_144", lambda: check_attention_bwd(B=2, heads=1, kept=(1, 1), Nq=144, Nkv=144)),
    ("bwd_scale_cols", check_scale_cols),
    ("bwd_scale_cols_wide", lambda: check_scale_cols(B=2, hw=64, C=3840, group=64)),
    ("bwd_geglu", check_geglu_train),
    ("bwd_geglu_nogate", lambda: check_geglu_train(use_gate=False, inner=1280)),
    ("bwd_groupnorm", check_groupnorm_bwd),
    ("bwd_groupnorm_plain_acc", lambda: check_groupnorm_bwd(silu=False, use_gate=False, accumulate=True)),
    ("bwd_groupnorm_wide", lambda: check_groupnorm_bwd(B=2, hw=64, C=2560, use_gate=False)),
    ("bwd_layernorm", check_layernorm_bwd),
    ("bwd_layernorm_320", lambda: check_layernorm_bwd(rows=100, C=320, accumulate=False)),
    ("bwd_layernorm_1280", lambda: check_layernorm_bwd(rows=64, C=1280)),
    ("bwd_depth_lerp", check_depth_lerp_bwd),
    ("bwd_resample", check_resample_bwd),
    ("bwd_conv_dgrad", check_conv_dgrad),
    ("bwd_conv_dgrad_s2", lambda: check_conv_dgrad(stride=2)),
]
