"""TEST INFRASTRUCTURE: a torch-on-CPU emulation of the C-ABI kernels' *contract* (segments, tiles,
pitches, column offsets, zero padding), monkey-patched over diffusion_pruning_b200.kernels so the
engine's host logic (expert bucketing, weight compaction, schedules, layer wiring) can be checked
against the oracle in the GPU-less build container. It is never imported by the product; the product
path has no fallback and raises without the CUDA library / a CUDA device."""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from diffusion_pruning_b200 import kernels as K
from diffusion_pruning_b200._lib import A_CONV3X3, A_CONV3X3_S2, A_LINEAR, EPI_GEGLU, EPI_SILU, OUT_BF16, OUT_F32, OUT_F32_NCHW


def _view(t: torch.Tensor, rows: int, cols: int, ld: int) -> torch.Tensor:
    return torch.as_strided(t, (rows, cols), (ld, 1), t.storage_offset())


def _check_tiles(sched: K.Schedule, mode, Ho, Wo, geglu):
    """Every active segment must be exactly covered by its tiles."""
    segs = sched.segs.numpy()
    tiles = sched.tiles.numpy()
    tiles = tiles[(tiles[:, 3] & 8) == 0]  # APTP_TILE_SKIP: padding entries of an A-stationary list
    bw, bh, bb = sched.box
    cols = sched.bn // 2 if geglu else sched.bn
    for si in range(sched.n_segs):
        rb, re, nv, ns = segs[si][:4]
        mine = tiles[tiles[:, 0] == si]
        if len(mine) == 0:
            continue
        nt = (max(nv, ns) + cols - 1) // cols
        w32 = int(mine[0, 3] >> 8) & 0xFF
        width = w32 * 32 if w32 else sched.bn  # balanced tiles: narrower than bn, same count
        assert width <= sched.bn and nt * (width // 2 if geglu else width) >= max(nv, ns)
        assert set(mine[:, 2].tolist()) == {i * width for i in range(nt)}
        covered = np.zeros(re - rb, dtype=np.int32)
        for m in np.unique(mine[:, 1]):
            if mode == A_LINEAR:
                r = np.arange(m, m + 128)
            else:
                hw = Ho * Wo
                img0, rem = divmod(int(m), hw)
                oy0, ox0 = divmod(rem, Wo)
                rr = np.arange(128)
                ix, iy, ib = rr % bw, (rr // bw) % bh, rr // (bw * bh)
                r = ((img0 + ib) * Ho + oy0 + iy) * Wo + ox0 + ix
            r = r[(r >= rb) & (r < re)]
            covered[r - rb] += 1
        assert (covered == 1).all(), "tiles must cover each segment row exactly once"


def grouped_gemm(a, w, out, sched, *, a_ld, a_k, a_rows, mode=A_LINEAR, batch=1, H=1, W=1, k_tap_pitch=0, out_ld,
                 out_mode=OUT_BF16, bias=None, rowvec=None, rowvec_ld=0, rows_per_sample=1, residual=None, res_ld=0,
                 gate=None, gate_ld=0, gate_group=1, border_tab=None, tab_ld=0, flags=0, ln_colsum=None, ln_rowstats=None,
                 rowstat_out=None, colstat=None, a2=None, a2_ld=0, a2_k=0, w2=None):
    if sched.n_tiles == 0:
        return
    geglu = bool(flags & EPI_GEGLU)
    stride = 2 if mode == A_CONV3X3_S2 else 1
    Ho, Wo = (H // stride, W // stride) if mode != A_LINEAR else (1, 1)
    _check_tiles(sched, mode, Ho, Wo, geglu)
    segs = sched.segs.numpy()
    tiles = sched.tiles.numpy()
    tiles = tiles[(tiles[:, 3] & 8) == 0]
    A = _view(a, a_rows, a_ld, a_ld).float()
    A[:, a_k:] = 0  # reads beyond the tensor-map extent are zero-filled
    Wm = w.float()
    for si in range(sched.n_segs):
        rb, re, nv, ns, kch, wro, voff, toff, oco = [int(v) for v in segs[si][:9]]
        if not (tiles[:, 0] == si).any():
            continue
        kk = kch * 64
        n_rows_w = (2 if geglu else 1) * 0  # placeholder
        if mode == A_LINEAR:
            Ae = A[rb:re, :kk] if kk <= a_ld else F.pad(A[rb:re], (0, kk - a_ld))
            if geglu:
                half = sched.bn // 2
                nt = (max(nv, ns) + half - 1) // half
                blk = Wm[wro:wro + nt * sched.bn, :kk].reshape(nt, 2, half, kk)
                wh = blk[:, 0].reshape(nt * half, kk)
                wg = blk[:, 1].reshape(nt * half, kk)
                acc_h, acc_g = Ae @ wh.t(), Ae @ wg.t()
                if ln_rowstats is not None:  # APTP_EPI_LN_FOLD: rstd * (acc - mean * colsum), per row
                    mu, rstd = ln_rowstats[rb:re, 0], ln_rowstats[rb:re, 1]
                    cs = ln_colsum[voff:voff + nt * sched.bn].reshape(nt, 2, half)
                    acc_h = rstd[:, None] * (acc_h - mu[:, None] * cs[:, 0].reshape(-1)[None])
                    acc_g = rstd[:, None] * (acc_g - mu[:, None] * cs[:, 1].reshape(-1)[None])
                if bias is not None:
                    bb_ = bias[voff:voff + nt * sched.bn].reshape(nt, 2, half)
                    acc_h = acc_h + bb_[:, 0].reshape(-1)
                    acc_g = acc_g + bb_[:, 1].reshape(-1)
                acc = None
            else:
                ncols = max(nv, ns)
                acc = Ae @ Wm[wro:wro + ncols, :kk].t()
        else:
            hw = Ho * Wo
            b0, b1 = rb // hw, re // hw
            x = A.reshape(batch, H, W, a_ld)[b0:b1, :, :, :kk].permute(0, 3, 1, 2)
            ncols = max(nv, ns)
            wk = Wm[wro:wro + ncols].reshape(ncols, 9, k_tap_pitch)[:, :, :kk].reshape(ncols, 3, 3, kk).permute(0, 3, 1, 2)
            acc = F.conv2d(x, wk, None, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(-1, ncols)
            if a2 is not None:  # second operand pair: 1x1 conv over a second tensor into the same tiles
                assert mode == A_CONV3X3
                A2 = _view(a2, a_rows, a2_k, a2_ld).float()
                acc = acc + A2[rb:re] @ w2.float()[:ncols, :a2_k].t()
        rows = torch.arange(rb, re)
        sample = rows // rows_per_sample
        if geglu:
            cols = torch.arange(acc_h.shape[1])
            if gate is not None:
                gsel = gate[sample][:, (cols // gate_group).clamp(max=gate_ld - 1)]
                acc_h, acc_g = acc_h * gsel, acc_g * gsel
            acc = acc_h * F.gelu(acc_g)
        else:
            cols = torch.arange(acc.shape[1])
            if ln_rowstats is not None:
                mu, rstd = ln_rowstats[rb:re, 0], ln_rowstats[rb:re, 1]
                nb = min(nv, acc.shape[1])
                acc[:, :nb] = rstd[:, None] * (acc[:, :nb] - mu[:, None] * ln_colsum[voff:voff + nb][None])
            if bias is not None:
                nb = min(nv, acc.shape[1])
                acc[:, :nb] += bias[voff:voff + nb]
            if rowvec is not None:
                rv = _view(rowvec, int(sample.max()) + 1, acc.shape[1], rowvec_ld)
                acc = acc + rv[sample]
            if border_tab is not None:
                pix = rows % (Ho * Wo)
                oy, ox = pix // Wo, pix % Wo
                yc = torch.where(oy == 0, 0, torch.where(oy == Ho - 1, 2, 1))
                xc = torch.where(ox == 0, 0, torch.where(ox == Wo - 1, 2, 1))
                tab = border_tab.reshape(-1)[toff:toff + 9 * tab_ld].reshape(9, tab_ld)[:, :acc.shape[1]]
                acc = acc + tab[yc * 3 + xc]
            if gate is not None:
                acc = acc * gate[sample][:, (cols // gate_group).clamp(max=gate_ld - 1)]
            if flags & EPI_SILU:
                acc = F.silu(acc)
        if residual is not None:
            R = _view(residual, re, res_ld, res_ld).float()
            nb = min(nv, acc.shape[1])
            acc[:, :nb] += R[rb:re, oco:oco + nb]
        acc[:, nv:] = 0
        nst = max(ns, nv) if out_mode != OUT_F32_NCHW else nv
        if out_mode == OUT_BF16:
            O = _view(out, re, out_ld, out_ld)
            O[rb:re, oco:oco + nst] = acc[:, :nst].to(torch.bfloat16)
            if colstat is not None:  # statistics of the STORED (bf16-rounded) values
                assert rows_per_sample % 128 == 0 and rb % rows_per_sample == 0
                a32 = acc[:, :nst].to(torch.bfloat16).float().reshape((re - rb) // 32, 32, nst)
                colstat[0][rb // 32: re // 32, oco:oco + nst] = a32.sum(1)
                colstat[1][rb // 32: re // 32, oco:oco + nst] = (a32 * a32).sum(1)
            if rowstat_out is not None:  # per-row (sum, sumsq) per 32-column chunk of the stored values
                assert oco % 32 == 0 and nst % 32 == 0
                a32 = acc[:, :nst].reshape(re - rb, nst // 32, 32)
                rowstat_out[rb:re, oco // 32: oco // 32 + nst // 32, 0] = a32.sum(-1)
                rowstat_out[rb:re, oco // 32: oco // 32 + nst // 32, 1] = (a32 * a32).sum(-1)
        elif out_mode == OUT_F32:
            O = _view(out, re, out_ld, out_ld)
            O[rb:re, oco:oco + nst] = acc[:, :nst]
            if colstat is not None:  # per-channel partial sums of every 32-row block (block order differs from the
                # device's tile / quadrant order for conv boxes; the consumer only sums over the blocks of a sample)
                assert rows_per_sample % 128 == 0 and rb % rows_per_sample == 0
                a32 = acc[:, :nst].reshape((re - rb) // 32, 32, nst)
                colstat[0][rb // 32: re // 32, oco:oco + nst] = a32.sum(1)
                colstat[1][rb // 32: re // 32, oco:oco + nst] = (a32 * a32).sum(1)
        else:
            nb = re // rows_per_sample
            O = torch.as_strided(out, (nb, out_ld, rows_per_sample), (out_ld * rows_per_sample, rows_per_sample, 1),
                                 out.storage_offset())
            O[rb // rows_per_sample:nb, :nv] = acc[:, :nv].reshape(-1, rows_per_sample, nv).permute(0, 2, 1)


def groupnorm_stats(x0, c0, ld0, x1, c1, ld1, batch, hw, group_size, sample_channels, stats, stats_groups,
                    x_f32=False):
    X = _view(x0, batch * hw, c0, ld0).float()
    if c1:
        X = torch.cat([X, _view(x1, batch * hw, c1, ld1).float()], 1)
    st = stats.view(batch, stats_groups, 2)
    for b in range(batch):
        ct = int(sample_channels[b]) if sample_channels is not None else c0 + c1
        if ct <= 0:
            continue
        xb = X[b * hw:(b + 1) * hw, :ct]
        g = (ct + group_size - 1) // group_size
        xg = xb.reshape(hw, g, group_size)
        st[b, :g, 0] = xg.sum((0, 2))  # overwritten, not accumulated (deterministic two-stage reduction)
        st[b, :g, 1] = (xg * xg).sum((0, 2))


def groupnorm_stats_from_partials(cs0, c0, cs1, c1, blocks, batch, group_size, sample_channels, stats, stats_groups):
    S = cs0[0].view(batch, blocks, -1)[:, :, :c0].sum(1)
    Q = cs0[1].view(batch, blocks, -1)[:, :, :c0].sum(1)
    if cs1 is not None:
        S = torch.cat([S, cs1[0].view(batch, blocks, -1)[:, :, :c1].sum(1)], 1)
        Q = torch.cat([Q, cs1[1].view(batch, blocks, -1)[:, :, :c1].sum(1)], 1)
    st = stats.view(batch, stats_groups, 2)
    for b in range(batch):
        ct = int(sample_channels[b]) if sample_channels is not None else S.shape[1]
        if ct <= 0:
            continue
        g = (ct + group_size - 1) // group_size
        st[b, :g, 0] = S[b, :ct].reshape(g, group_size).sum(1)
        st[b, :g, 1] = Q[b, :ct].reshape(g, group_size).sum(1)


def groupnorm_apply(x0, c0, ld0, x1, c1, ld1, y, ldy, batch, hw, group_size, eps, stats, stats_groups, gamma, beta,
                    affine_ld, sample_seg, sample_channels, gate, gate_ld, silu, x_f32=False, raw_out=None, raw_ld=0):
    X = _view(x0, batch * hw, c0, ld0).float()
    if c1:
        X = torch.cat([X, _view(x1, batch * hw, c1, ld1).float()], 1)
    if raw_out is not None:
        R = _view(raw_out, batch * hw, raw_ld, raw_ld)
        for b in range(batch):
            ct = int(sample_channels[b]) if sample_channels is not None else c0 + c1
            if ct > 0:
                R[b * hw:(b + 1) * hw, :ct] = X[b * hw:(b + 1) * hw, :ct].to(torch.bfloat16)
    Y = _view(y, batch * hw, ldy, ldy)
    st = stats.view(batch, stats_groups, 2)
    G = gamma.reshape(-1, affine_ld)
    Bt = beta.reshape(-1, affine_ld)
    for b in range(batch):
        ct = int(sample_channels[b]) if sample_channels is not None else c0 + c1
        if ct <= 0:
            continue
        seg = int(sample_seg[b]) if sample_seg is not None else 0
        g = ct // group_size
        n = hw * group_size
        mean = st[b, :g, 0] / n
        var = (st[b, :g, 1] / n - mean * mean).clamp(min=0)
        gt = gate[b, :g] if gate is not None else torch.ones(g)
        rstd = torch.rsqrt(gt * gt * var + eps)
        w = G[seg, :ct] * (rstd * gt).repeat_interleave(group_size)
        sh = Bt[seg, :ct] - mean.repeat_interleave(group_size) * w
        o = X[b * hw:(b + 1) * hw, :ct] * w + sh
        if silu:
            o = F.silu(o)
        cs = min((ct + 63) // 64 * 64, ldy)
        Y[b * hw:(b + 1) * hw, :ct] = o.to(torch.bfloat16)
        Y[b * hw:(b + 1) * hw, ct:cs] = 0


def ln_rowstats(partial, rows, C_, eps, out, sample_active=None, rows_per_sample=1):
    ps = partial[:rows].sum(1)
    mu = ps[:, 0] / C_
    out[:rows, 0] = mu
    out[:rows, 1] = torch.rsqrt((ps[:, 1] / C_ - mu * mu).clamp(min=0) + eps)


def layernorm(x, ldx, y, ldy, rows, C_, eps, gamma, beta, sample_active=None, rows_per_sample=1):
    X = _view(x, rows, C_, ldx).float()
    Y = _view(y, rows, C_, ldy)
    o = F.layer_norm(X, (C_,), gamma, beta, eps).to(torch.bfloat16)
    if sample_active is not None:
        act = sample_active.bool().repeat_interleave(rows_per_sample)
        Y[act] = o[act]
    else:
        Y.copy_(o)


def depth_lerp(x, ldx, y, ldy, out, ldo, rows, C_, d, rows_per_sample):
    X = _view(x, rows, C_, ldx).float()
    Yv = _view(y, rows, C_, ldy).float()
    dd = d.repeat_interleave(rows_per_sample)[:, None]
    _view(out, rows, C_, ldo).copy_(((1 - dd) * X + dd * Yv).to(torch.bfloat16))


def copy_rows(src, lds, dst, ldd, rows, C_, sample_mask=None, rows_per_sample=1):
    S = _view(src, rows, C_, lds)
    D = _view(dst, rows, C_, ldd)
    if sample_mask is not None:
        m = sample_mask.bool().repeat_interleave(rows_per_sample)
        D[m] = S[m]
    else:
        D.copy_(S)


def copy_rows_cvt(src, lds, dst, ldd, rows, C_, sample_mask=None, rows_per_sample=1):
    S = _view(src, rows, C_, lds).to(dst.dtype)
    D = _view(dst, rows, C_, ldd)
    if sample_mask is not None:
        m = sample_mask.bool().repeat_interleave(rows_per_sample)
        D[m] = S[m]
    else:
        D.copy_(S)


def depth_lerp_f32(x, ldx, y, ldy, out, ldo, rows, C_, d, rows_per_sample):
    X = _view(x, rows, C_, ldx)
    Yv = _view(y, rows, C_, ldy).clone()
    dd = d.repeat_interleave(rows_per_sample)[:, None]
    _view(out, rows, C_, ldo).copy_((1 - dd) * X + dd * Yv)


def upsample2x_cvt(src, dst, batch, H, W, C_):
    upsample2x(src.to(torch.bfloat16), dst, batch, H, W, C_)


def upsample2x(src, dst, batch, H, W, C_):
    s = src.view(batch, H, W, C_)
    dst.view(batch, 2 * H, 2 * W, C_).copy_(s.repeat_interleave(2, 1).repeat_interleave(2, 2))


def im2col_input(sample_nchw, dst, batch, cin, H, W):
    cols = F.unfold(sample_nchw, 3, padding=1).reshape(batch, cin, 9, H * W).permute(0, 3, 2, 1).reshape(batch * H * W, 9 * cin)
    dst.zero_()
    dst[:, :9 * cin] = cols.to(torch.bfloat16)


def timestep_embedding(t, dst, batch, dim):
    half = dim // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = t[:, None] * f[None]
    dst.copy_(torch.cat([torch.cos(arg), torch.sin(arg)], 1).to(torch.bfloat16))


def cast_f32_bf16(src, dst, n):
    dst.view(-1)[:n] = src.reshape(-1)[:n].to(torch.bfloat16)


def silu_bf16(src, dst, n):
    dst.view(-1)[:n] = F.silu(src.reshape(-1)[:n].float()).to(torch.bfloat16)


def attention(q, ldq, k, ldk, v, ldv, out, ldo, batch, n_q, n_kv, sample_heads, max_heads, scale, lse2=None):
    for b in range(batch):
        nh = int(sample_heads[b])
        if nh == 0:
            continue
        Q = _view(q, (b + 1) * n_q, nh * 64, ldq)[b * n_q:].float().reshape(n_q, nh, 64).transpose(0, 1)
        Kk = _view(k, (b + 1) * n_kv, nh * 64, ldk)[b * n_kv:].float().reshape(n_kv, nh, 64).transpose(0, 1)
        V = _view(v, (b + 1) * n_kv, nh * 64, ldv)[b * n_kv:].float().reshape(n_kv, nh, 64).transpose(0, 1)
        p = torch.softmax(Q @ Kk.transpose(1, 2) * scale, -1)
        o = (p @ V).transpose(0, 1).reshape(n_q, nh * 64)
        _view(out, (b + 1) * n_q, nh * 64, ldo)[b * n_q:] = o.to(torch.bfloat16)


def check_abort():
    return None


def install(monkeypatch):
    for name in ("grouped_gemm", "groupnorm_stats", "groupnorm_stats_from_partials", "groupnorm_apply", "layernorm", "ln_rowstats", "depth_lerp", "copy_rows",
                 "copy_rows_cvt", "depth_lerp_f32", "upsample2x_cvt",
                 "upsample2x", "im2col_input", "timestep_embedding", "cast_f32_bf16", "silu_bf16", "attention",
                 "check_abort"):
        monkeypatch.setattr(K, name, globals()[name])


# ------------------------------------------------------------------------------------------------
# router kernels (contract of csrc/router.cu), so the quantizer's HOST logic -- uniform draw order, phase
# protocol of the sharded Sinkhorn with its all-reduces -- can run under gloo in the GPU-less container
# ------------------------------------------------------------------------------------------------
def gumbel_gate(z, u, out, batch, n_width, n_depth, width_starts, n_gates, depth_order, temperature, base,
                non_zero_width):
    g = -torch.log(-torch.log(u + 1e-20) + 1e-20)
    y = torch.sigmoid((z[:, :n_width] + g[:, :n_width] + base) / temperature)
    ws = width_starts.tolist()
    if non_zero_width:
        y = y.clone()
        for i in range(n_gates):
            s, e = ws[i], ws[i + 1]
            dead = (y[:, s:e] >= 0.5).sum(1) == 0
            y[dead, s] += 0.5
    out[:, :n_width] = y
    if n_depth:
        x = torch.flip(torch.cumsum(torch.softmax(z[:, n_width:], 1), 1), dims=[1])
        x = torch.log(x + 1e-6) - torch.log1p(-(x - 1e-6))
        yd = torch.sigmoid((x + g[:, n_width:] + base) / temperature)
        out[:, n_width + depth_order.long()] = yd


def arch_normalize(gates, out, batch, dim, col_depth, col_scale, l2=True):
    cd = col_depth.long()
    hard = (gates >= 0.5).float()
    dep = gates[:, cd.clamp_min(0)]
    v = torch.where(cd[None, :] >= 0, gates * dep, hard) * col_scale[None, :]
    out.copy_(v / v.norm(dim=1, keepdim=True) if l2 else v)


def route_cosine(a_norm, codes_norm, scores, indices, batch, dim, n_codes):
    s = (a_norm.double() @ codes_norm.double().t()).float()
    scores.copy_(s)
    indices.copy_(torch.argmax(s, dim=1))


def sinkhorn_phase(phase, Q, scores, partial, indices, batch_local, batch_global, n_codes, epsilon, first_iter):
    if phase == 0:
        Q.copy_(torch.exp(scores / epsilon))
        partial[0] = Q.double().sum()
    elif phase == 1:
        if first_iter:
            Q.div_(partial[0].float())
        partial[1:1 + n_codes] = Q.double().sum(0)
    elif phase == 2:
        Q.div_(partial[1:1 + n_codes].float()[None, :])
        Q.div_(float(n_codes))
        Q.div_(Q.sum(1, keepdim=True))
        Q.div_(float(batch_global))
    else:
        Q.mul_(float(batch_global))
        indices.copy_(torch.argmax(Q, dim=1))


def route_sinkhorn(scores, Q, partial, indices, batch, n_codes, epsilon, iterations):
    sinkhorn_phase(0, Q, scores, partial, indices, batch, batch, n_codes, epsilon, 0)
    for it in range(iterations):
        sinkhorn_phase(1, Q, scores, partial, indices, batch, batch, n_codes, epsilon, it == 0)
        sinkhorn_phase(2, Q, scores, partial, indices, batch, batch, n_codes, epsilon, 0)
    sinkhorn_phase(3, Q, scores, partial, indices, batch, batch, n_codes, epsilon, 0)


def install_router(setattr_fn):
    """setattr_fn(obj, name, value): monkeypatch.setattr or plain setattr (spawned workers)."""
    from diffusion_pruning_b200.quantizer import StructureVectorQuantizer
    for name in ("gumbel_gate", "arch_normalize", "route_cosine", "sinkhorn_phase", "route_sinkhorn"):
        setattr_fn(K, name, globals()[name])
    setattr_fn(StructureVectorQuantizer, "_require_cuda", staticmethod(lambda t, what: None))
