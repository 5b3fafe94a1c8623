"""Training-mode (soft-gate, differentiable) execution of the gated U-Net: the student forward of the
pruning step (pdm/training/trainer.py:1192-1195) and its backward w.r.t. the 84 gate tensors.

All U-Net weights are frozen during pruning (pdm/models/unet/unet_2d_conditional.py:2118-2122) and the
inputs are detached, so the backward produces only activation gradients (bf16) and per-(sample, gate)
reductions (fp32): dgrad GEMMs / convs run on the same tcgen05 grouped GEMM with transposed (conv:
tap-flipped) weights, attention backward on its two tcgen05 kernels, and the gated elementwise /
normalisation ops on the K5 kernels of csrc/backward.cu. There is no autograd graph
inside the U-Net: the forward records a tape, `UNetTrainFunction` exposes the whole step to PyTorch as
ONE autograd node whose differentiable inputs are the gate tensors handed to `set_structure` and whose
outputs are the prediction plus the nine hooked block activations (trainer.py:496-511, :1220-1225).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import kernels as K
from . import plan as P
from ._lib import A_CONV3X3, A_LINEAR, EPI_SILU, OUT_BF16, OUT_F32, OUT_F32_NCHW
from .unet import BF16, Act, ResnetBlock2DWidthGated, Transformer2DModelWidthGated, _Engine


class _G:
    """A gradient in engine layout: bf16 rows [M, C] behind an explicit pitch (may be a column slice)."""
    __slots__ = ("t", "ld", "C")

    def __init__(self, t: torch.Tensor, ld: int, C: int):
        self.t, self.ld, self.C = t, ld, C


class TrainEngine(_Engine):
    """Soft-gate forward with a tape + explicit backward. Sample order is the caller's (no bucketing:
    with soft gates nothing can be skipped)."""

    # ------------------------------------------------------------------------------------------
    # small helpers
    # ------------------------------------------------------------------------------------------
    def _new(self, rows: int, cols: int, dtype=BF16, zero: bool = False) -> torch.Tensor:
        f = torch.zeros if zero else torch.empty
        return f(rows, cols, device=self.device, dtype=dtype)

    # ---- weight gradients (fine-tune stage: every U-Net parameter trains, trainer.py:1683-1765) ----------
    train_weights = False  # set by UNetFineTuneFunction; the pruning stage keeps the U-Net frozen

    def _wg_buf(self, param: torch.Tensor, shape) -> torch.Tensor:
        """fp32 accumulator of one parameter's gradient in KERNEL layout (OHWI for convs, [C, 2] = (gamma, beta) pairs
        are handled by _affine): zero-filled once per step, the kernels accumulate into it."""
        g = self.wg.get(id(param))
        if g is None:
            g = torch.zeros(*shape, device=self.device, dtype=torch.float32)
            self.wg[id(param)] = g
        return g

    def _wgrad(self, weight, bias, dy: torch.Tensor, ld_dy: int, a: torch.Tensor, ld_a: int, rows: int, n_out: int,
               k_in: int, conv=None, stride: int = 1):
        if not self.train_weights:
            return
        taps = 9 if conv is not None else 1
        dw = self._wg_buf(weight, (n_out, taps * k_in))
        db = self._wg_buf(bias, (n_out,)) if bias is not None else None
        K.wgrad(dy, ld_dy, a, ld_a, dw, db, rows, n_out, k_in, conv=conv, stride=stride)
        self.launches += 1

    def _affine(self, norm: nn.Module) -> Optional[torch.Tensor]:
        """[C, 2] accumulator of (dgamma, dbeta) of a GroupNorm / LayerNorm, or None when the weights are frozen."""
        if not self.train_weights:
            return None
        return self._wg_buf(norm.weight, (norm.weight.shape[0], 2))

    def _mm(self, key, a: torch.Tensor, a_ld: int, a_k: int, M: int, w: torch.Tensor, N: int, out: torch.Tensor,
            out_ld: int, *, bias=None, residual=None, res_ld=0, rows_per_sample=1, flags=0, out_col_off=0,
            conv: Optional[Tuple[int, int, int]] = None, k_tap_pitch=0, rowvec=None, rowvec_ld=0, out_mode=OUT_BF16):
        """Single-segment dense GEMM / 3x3 conv (stride 1) on the grouped tcgen05 kernel."""
        mode = A_LINEAR if conv is None else A_CONV3X3
        sk = ("tr", key, M, N, a_k, conv, out_col_off)
        sched = self.sched.get(sk)
        if sched is None:
            seg = K.Segment(0, M, N, (a_k + 63) // 64, out_col_off=out_col_off)
            bn = P.choose_bn([max(N, 32)])
            if conv is None:
                sched = K.build_schedule([seg], bn, self.device)
            else:
                sched = K.build_schedule([seg], bn, self.device, mode=A_CONV3X3, Ho=conv[1], Wo=conv[2])
            self.sched[sk] = sched
        kw: Dict[str, Any] = {}
        if conv is not None:
            kw = dict(mode=A_CONV3X3, batch=conv[0], H=conv[1], W=conv[2], k_tap_pitch=k_tap_pitch)
        K.grouped_gemm(a, w, out, sched, a_ld=a_ld, a_k=a_k, a_rows=M, out_ld=out_ld, out_mode=out_mode, bias=bias,
                       rows_per_sample=rows_per_sample, residual=residual, res_ld=res_ld, flags=flags, rowvec=rowvec,
                       rowvec_ld=rowvec_ld, **kw)
        self.launches += 1
        self.flops += sched.flops

    def _wT(self, name: str, mod: nn.Module, pad_out_to: int = 0) -> torch.Tensor:
        """Transposed weight for the dgrad: Linear / 1x1 conv -> [K, N]; 3x3 conv -> [Cin, 9*Cout] tap-flipped."""
        def build():
            w = mod.weight.detach().to(self.device)
            if w.ndim == 4 and w.shape[-1] == 3:
                if pad_out_to and w.shape[0] < pad_out_to:
                    w = torch.cat([w, torch.zeros(pad_out_to - w.shape[0], *w.shape[1:], device=self.device,
                                                  dtype=w.dtype)], 0)
                return P.pack_conv_weight_dgrad(w).to(BF16).contiguous()
            return w.reshape(w.shape[0], -1).t().to(BF16).contiguous()
        return self._pack(self.dense, name + ".T", build)

    def _w(self, name: str, mod: nn.Module) -> Dict[str, torch.Tensor]:
        return self._dense_linear(name, mod)

    def _gn_again(self, x: torch.Tensor, ld: int, C: int, B: int, hw: int, groups: int, eps: float, gamma, beta,
                  out: torch.Tensor, silu: bool, stats: torch.Tensor, gate=None):
        """Recompute a GroupNorm(+gate, +SiLU) output from the statistics saved on the tape (one HBM pass): the
        activations feeding a conv / linear are not kept for the backward, only the block inputs are."""
        K.groupnorm_apply(x, C, ld, None, 0, 0, out, C, B, hw, C // groups, eps, stats, groups, gamma, beta, C, None, None,
                          gate, gate.stride(0) if gate is not None else groups, silu)
        self.launches += 1

    def _gn_fwd(self, x: torch.Tensor, ld: int, C: int, B: int, hw: int, groups: int, eps: float, gamma, beta,
                out: torch.Tensor, silu: bool, gate=None) -> torch.Tensor:
        stats = self._take_stats(B, groups)
        gs = C // groups
        K.groupnorm_stats(x, C, ld, None, 0, 0, B, hw, gs, None, stats, groups)
        K.groupnorm_apply(x, C, ld, None, 0, 0, out, C, B, hw, gs, eps, stats, groups, gamma, beta, C, None, None, gate,
                          gate.stride(0) if gate is not None else groups, silu)
        self.launches += 2
        return stats

    def _take_stats(self, B: int, groups: int) -> torch.Tensor:
        """[B, groups, 2] zeroed fp32 slice of this step's statistics pool (ONE fill per step instead of one per
        GroupNorm; the slices live on the tape until the backward has consumed them)."""
        n = B * groups * 2
        if self._stats_pool is None or self._stats_used + n > self._stats_pool.numel():
            self._stats_pool = torch.zeros(max(96 * n, 1 << 16), device=self.device, dtype=torch.float32)
            self._stats_used = 0
        v = self._stats_pool[self._stats_used:self._stats_used + n].view(B, groups, 2)
        self._stats_used += n
        return v

    def _dgate(self, gate_idx: int) -> torch.Tensor:
        """Gradient slot of one width gate: a VIEW into the [B, 1620] accumulator (same row pitch as the gate
        matrix; the backward kernels accumulate into it atomically)."""
        s, e = self.width_starts[gate_idx], self.width_starts[gate_idx + 1]
        return self.darch[:, s:e]

    def _gn_bwd(self, x, ld, C, B, hw, groups, eps, stats, gamma, beta, da, dx: torch.Tensor, lddx: int,
                accumulate: bool, silu: bool, gate=None, dgate=None, daffine=None):
        bstats = self.buf("gn_bstats", B, groups * 2, torch.float32)
        gld = gate.stride(0) if gate is not None else groups
        if daffine is not None:
            K.groupnorm_bwd_affine(x, ld, da, C, dx, lddx, accumulate, B, hw, C, C // groups, eps, stats, groups, gamma,
                                   beta, gate, gld, silu, bstats, dgate, daffine)
        else:
            K.groupnorm_bwd(x, ld, da, C, dx, lddx, accumulate, B, hw, C, C // groups, eps, stats, groups, gamma, beta,
                            gate, gld, silu, bstats, dgate)
        self.launches += 2

    # ------------------------------------------------------------------------------------------
    # forward with tape
    # ------------------------------------------------------------------------------------------
    def run_train(self, sample: torch.Tensor, timestep, ctx: torch.Tensor, gate_list: List[torch.Tensor]):
        m = self.m
        B, cin, H, W = sample.shape
        self.B = B
        self.flops, self.launches = 0.0, 0
        self.compact, self.eset, self.layout = False, None, None
        # the tape holds ONE forward: stamp it so a backward that belongs to an older forward is refused, not silently
        # run on the newer tape (two grad-enabled forwards before one backward)
        self.tape_id = getattr(self, "tape_id", 0) + 1
        # gates: [B, 1620] fp32 in get_structure order (widths then depths), as the caller supplied them
        widths = [w for ws in m.get_structure()["width"] for w in ws]
        n_w = len(widths)
        arch = torch.cat([g.detach().reshape(g.shape[0], -1).float() for g in gate_list[:n_w]] +
                         [g.detach().reshape(-1, 1).float() for g in gate_list[n_w:]], dim=1).to(self.device)
        if arch.shape[0] != B:
            assert B % arch.shape[0] == 0
            arch = arch.repeat(B // arch.shape[0], 1)
        self.soft_arch = arch.contiguous()
        self.gate_rows = gate_list[0].shape[0]
        self.width_starts = [0]
        for w in widths:
            self.width_starts.append(self.width_starts[-1] + w)
        self.n_width = self.width_starts[-1]
        self.gate_cols = {}
        gi = di = 0
        for mod in m._gated:
            n = len(mod.gate_widths())
            self.gate_cols[mod.uid] = {"w": list(range(gi, gi + n)), "d": (di if mod.depth_gate is not None else None)}
            gi += n
            di += int(mod.depth_gate is not None)
        self.darch = torch.zeros_like(self.soft_arch)  # gradient accumulator, same column layout and row pitch
        assert self.darch.stride(0) == self.soft_arch.stride(0)
        self.ddepth_T = torch.zeros(self.soft_arch.shape[1] - self.n_width, B, device=self.device,
                                    dtype=torch.float32)  # [14, B] depth-gate gradients (contiguous per gate)
        self._stats_pool, self._stats_used = None, 0
        self.tape: List[Tuple] = []
        self.first_uid = m._gated[0].uid  # nothing upstream of it needs a gradient

        if not torch.is_tensor(timestep):
            timestep = torch.tensor([float(timestep)], device=self.device)
        sample = sample.detach().to(self.device, torch.float32).contiguous()
        ctx = ctx.detach().to(self.device).contiguous()
        self.wg: Dict[int, torch.Tensor] = {}
        self.param_grads: Dict[int, torch.Tensor] = {}
        self.n_ctx = ctx.shape[1]
        cdim = ctx.shape[2]
        ctx16 = self.buf("ctx", B * self.n_ctx, cdim)
        if ctx.dtype == BF16:
            ctx16.copy_(ctx.reshape(B * self.n_ctx, cdim))
        else:
            K.cast_f32_bf16(ctx.to(torch.float32), ctx16, B * self.n_ctx * cdim)
        self.ctx = ctx16
        self.time_embed(timestep)
        c0 = m.config["block_out_channels"][0]
        if self.train_weights:
            # the time-embedding MLP stores only its activated outputs (SiLU fused into the epilogues); the weight
            # backward needs the pre-activations: two more [B, 1280] GEMMs without the SiLU, fp32 out
            tdim = c0 * 4
            te_m = m.time_embedding
            d1, d2 = self._dense_linear("te1", te_m.linear_1), self._dense_linear("te2", te_m.linear_2)
            self.te_z1 = torch.empty(B, tdim, device=self.device, dtype=torch.float32)
            self.te_z2 = torch.empty(B, tdim, device=self.device, dtype=torch.float32)
            self._mm(("te1z",), self.buf("t_sin", B, c0), c0, c0, B, d1["w"], tdim, self.te_z1, tdim, bias=d1["b"],
                     out_mode=OUT_F32)
            self._mm(("te2z",), self.buf("t_h", B, tdim), tdim, tdim, B, d2["w"], tdim, self.te_z2, tdim, bias=d2["b"],
                     out_mode=OUT_F32)
            self.dproj_all = torch.zeros(B, m._temb_total, device=self.device, dtype=torch.float32)
        col = self.buf("im2col", B * H * W, 64)
        K.im2col_input(sample, col, B, cin, H, W)
        def build_conv_in():
            w = m.conv_in.weight.detach().to(self.device)
            wp = torch.zeros(c0, 64, device=self.device, dtype=BF16)
            wp[:, :9 * cin] = w.permute(0, 2, 3, 1).reshape(c0, 9 * cin).to(BF16)
            return {"w": wp, "b": m.conv_in.bias.detach().to(self.device, torch.float32).contiguous()}
        d = self._pack(self.dense, "conv_in", build_conv_in)
        x0 = self._new(B * H * W, c0)
        self._mm("conv_in", col, 64, 64, B * H * W, d["w"], c0, x0, c0, bias=d["b"], rows_per_sample=H * W)
        x = Act(x0, B, H, W, c0)
        self.x0_act, self.im2col = x, col
        taps: List[Act] = []
        skips = [x]
        for blk in m.down_blocks:
            for i, r in enumerate(blk.resnets):
                x = self.t_resnet(r, x)
                if blk.attentions is not None:
                    x = self.t_transformer(blk.attentions[i], x)
                skips.append(x)
            if blk.downsamplers is not None:
                x = self.t_downsample(blk.downsamplers[0], x)
                skips.append(x)
            taps.append(x)
        mb = m.mid_block
        x = self.t_resnet(mb.resnets[0], x)
        x = self.t_transformer(mb.attentions[0], x)
        x = self.t_resnet(mb.resnets[1], x)
        taps.append(x)
        for blk in m.up_blocks:
            for i, r in enumerate(blk.resnets):
                x = self.t_resnet(r, x, skip=skips.pop())
                if blk.attentions is not None:
                    x = self.t_transformer(blk.attentions[i], x)
            if blk.upsamplers is not None:
                x = self.t_upsample(blk.upsamplers[0], x)
            taps.append(x)
        # conv_norm_out + SiLU + conv_out
        dn = self.dense.get("out_norm")
        if dn is None:
            dn = {"g": m.conv_norm_out.weight.detach().to(self.device, torch.float32).contiguous(),
                  "b": m.conv_norm_out.bias.detach().to(self.device, torch.float32).contiguous()}
            self.dense["out_norm"] = dn
        groups = m.config["norm_num_groups"]
        a = self.buf("gn_a", x.rows, x.C)
        stats = self._gn_fwd(x.t, x.ld, x.C, B, x.hw, groups, m.config["norm_eps"], dn["g"], dn["b"], a, True)
        cout = m.config["out_channels"]
        y = torch.empty(B, cout, x.H, x.W, device=self.device, dtype=torch.float32)
        dco = self._dense_linear("conv_out", m.conv_out, n_pad_to=32)
        sk = ("conv_out_tr", x.H, x.W, B)
        sched = self.sched.get(sk)
        if sched is None:
            sched = K.build_schedule([K.Segment(0, x.rows, cout, (x.C + 63) // 64)], 32, self.device, mode=A_CONV3X3,
                                     Ho=x.H, Wo=x.W)
            self.sched[sk] = sched
        K.grouped_gemm(a, dco["w"], y, sched, a_ld=x.C, a_k=x.C, a_rows=x.rows, mode=A_CONV3X3, batch=B, H=x.H, W=x.W,
                       k_tap_pitch=x.C, out_ld=cout, out_mode=OUT_F32_NCHW, bias=dco["b"], rows_per_sample=x.hw)
        self.final = (x, stats, dn)
        self.tap_acts = taps
        return y, taps

    # ---- ResNet ----------------------------------------------------------------------------------
    def t_resnet(self, r: ResnetBlock2DWidthGated, x: Act, skip: Optional[Act] = None) -> Act:
        B, H, W, hw, M = x.B, x.H, x.W, x.hw, x.rows
        if skip is not None:
            cat = self._new(M, x.C + skip.C)
            K.copy_rows(x.t, x.ld, cat, x.C + skip.C, M, x.C)
            K.copy_rows(skip.t, skip.ld, cat[:, x.C:], x.C + skip.C, M, skip.C)
            xin = Act(cat, B, H, W, x.C + skip.C)
        else:
            xin = x
        assert xin.C == r.cin
        pk = self._resnet_pack(r)
        cidx = self.gate_cols[r.uid]
        gate = self._soft_gate(cidx["w"][0])
        a1 = self.buf("gn_a", M, r.cin)
        st1 = self._gn_fwd(xin.t, xin.ld, r.cin, B, hw, r.groups, r.eps, pk["g1"], pk["b1"], a1, True)
        h1 = self._new(M, r.cout)
        rv = self.temb_rowvec[:, self.m._temb_off[r.uid]:]
        self._mm(("c1", r.uid), a1, r.cin, r.cin, M, pk["w1"], r.cout, h1, r.cout, conv=(B, H, W), k_tap_pitch=r.cin,
                 rowvec=rv, rowvec_ld=self.m._temb_total, rows_per_sample=hw)
        a2 = self.buf("gn_b", M, r.cout)
        st2 = self._gn_fwd(h1, r.cout, r.cout, B, hw, r.groups, r.eps, pk["gamma2"], pk["beta2"], a2, True, gate=gate)
        y = self._new(M, r.cout)
        if r.conv_shortcut is not None:
            sc = self._w("sc." + r.uid, r.conv_shortcut)
            self._mm(("sc", r.uid), xin.t, xin.ld, r.cin, M, sc["w"], r.cout, y, r.cout, bias=sc["b"], rows_per_sample=hw)
            res, res_ld = y, r.cout
        else:
            res, res_ld = xin.t, xin.ld
        self._mm(("c2", r.uid), a2, r.cout, r.cout, M, pk["w2"], r.cout, y, r.cout, conv=(B, H, W), k_tap_pitch=r.cout,
                 bias=pk["b2"], residual=res, res_ld=res_ld, rows_per_sample=hw)
        out, dgate_d = y, None
        keep_c = xin.C - (r.skip_connection_dim or 0)
        if r.depth_gate is not None:
            out = self._new(M, r.cout)
            dgate_d = self._soft_depth(cidx["d"])
            K.depth_lerp(xin.t, xin.ld, y, r.cout, out, r.cout, M, keep_c, dgate_d, hw)
        o = Act(out, B, H, W, r.cout)
        self.tape.append(("res", r, dict(x=x, skip=skip, xin=xin, st1=st1, h1=h1, st2=st2, y=y, gate=gate, d=dgate_d,
                                         keep=keep_c, out=o)))
        return o

    def b_resnet(self, r: ResnetBlock2DWidthGated, s: Dict[str, Any], dout: _G):
        xin: Act = s["xin"]
        B, H, W, hw, M = xin.B, xin.H, xin.W, xin.hw, xin.rows
        pk = self._resnet_pack(r)
        cidx = self.gate_cols[r.uid]
        need_in = r.uid != self.first_uid or self.train_weights  # norm1 / conv1 gradients need the input-side pass
        dxin = None
        if r.depth_gate is not None:
            dxin = self._new(M, r.cin, zero=(s["keep"] != r.cin))
            dy = self._new(M, r.cout)
            K.depth_lerp_bwd(dout.t, dout.ld, xin.t, xin.ld, s["y"], r.cout, dy, r.cout, dxin, r.cin, False, B, hw,
                             s["keep"], s["d"], self.ddepth_T[cidx["d"]])
            dyg = _G(dy, r.cout, r.cout)
            have = True
        else:
            dyg = dout
            have = False
        if need_in:
            if r.conv_shortcut is not None:
                wt = self._wT("sc." + r.uid, r.conv_shortcut)
                if dxin is None:
                    dxin = self._new(M, r.cin)
                self._mm(("sc.T", r.uid), dyg.t, dyg.ld, r.cout, M, wt, r.cin, dxin, r.cin,
                         residual=dxin if have else None, res_ld=r.cin, rows_per_sample=hw)
                self._wgrad(r.conv_shortcut.weight, r.conv_shortcut.bias, dyg.t, dyg.ld, xin.t, xin.ld, M, r.cout, r.cin)
            else:
                if have:
                    K.add_rows(dyg.t, dyg.ld, dxin, r.cin, M, r.cout)
                elif dyg.ld == r.cin and dyg.C == r.cin:
                    dxin = dyg.t  # identity skip: take ownership of the incoming gradient buffer
                else:
                    dxin = self._new(M, r.cin)
                    K.copy_rows(dyg.t, dyg.ld, dxin, r.cin, M, r.cin)
        # conv2 dgrad -> GN2(+gate,+SiLU) backward
        da2 = self.buf("bw_a", M, r.cout)
        self._mm(("c2.T", r.uid), dyg.t, dyg.ld, r.cout, M, self._wT("c2." + r.uid, r.conv2), r.cout, da2, r.cout,
                 conv=(B, H, W), k_tap_pitch=r.cout, rows_per_sample=hw)
        if self.train_weights:  # conv2: its input silu(GN2(gate * h1)) is recomputed from the saved statistics
            a2 = self.buf("gn_b", M, r.cout)
            self._gn_again(s["h1"], r.cout, r.cout, B, hw, r.groups, r.eps, pk["gamma2"], pk["beta2"], a2, True, s["st2"],
                           gate=s["gate"])
            self._wgrad(r.conv2.weight, r.conv2.bias, dyg.t, dyg.ld, a2, r.cout, M, r.cout, r.cout, conv=(B, H, W))
        dh1 = self.buf("bw_b", M, r.cout)
        self._gn_bwd(s["h1"], r.cout, r.cout, B, hw, r.groups, r.eps, s["st2"], pk["gamma2"], pk["beta2"], da2, dh1,
                     r.cout, False, True, gate=s["gate"], dgate=self._dgate(cidx["w"][0]), daffine=self._affine(r.norm2))
        if not need_in:
            return None
        if self.train_weights:
            a1 = self.buf("gn_a", M, r.cin)
            self._gn_again(xin.t, xin.ld, r.cin, B, hw, r.groups, r.eps, pk["g1"], pk["b1"], a1, True, s["st1"])
            self._wgrad(r.conv1.weight, None, dh1, r.cout, a1, r.cin, M, r.cout, r.cin, conv=(B, H, W))
            # h1 = conv1(a1) + (time_emb_proj(silu(temb)) + both biases)[b] broadcast over the sample's pixels
            off = self.m._temb_off[r.uid]
            K.col_sum_groups(dh1, r.cout, B, hw, r.cout, self.dproj_all[:, off:off + r.cout])
        da1 = self.buf("bw_c", M, r.cin)
        self._mm(("c1.T", r.uid), dh1, r.cout, r.cout, M, self._wT("c1." + r.uid, r.conv1), r.cin, da1, r.cin,
                 conv=(B, H, W), k_tap_pitch=r.cout, rows_per_sample=hw)
        if dxin is None:  # first resnet of the net when only the weights need this pass
            dxin = self._new(M, r.cin)
            self._gn_bwd(xin.t, xin.ld, r.cin, B, hw, r.groups, r.eps, s["st1"], pk["g1"], pk["b1"], da1, dxin, r.cin,
                         False, True, daffine=self._affine(r.norm1))
        else:
            self._gn_bwd(xin.t, xin.ld, r.cin, B, hw, r.groups, r.eps, s["st1"], pk["g1"], pk["b1"], da1, dxin, r.cin,
                         True, True, daffine=self._affine(r.norm1))
        return _G(dxin, r.cin, r.cin)

    # ---- transformer -----------------------------------------------------------------------------
    def _t_attn(self, uid: str, attn, gate_idx: int, xn: torch.Tensor, tok_in: torch.Tensor, tok_out: torch.Tensor,
                B: int, hw: int, C: int, cross: bool) -> Dict[str, Any]:
        M = B * hw
        heads = attn.heads
        gate = self._soft_gate(gate_idx)
        n_kv = self.n_ctx if cross else hw
        Mkv = B * n_kv
        wq, wk, wv = self._w("q." + uid, attn.to_q), self._w("k." + uid, attn.to_k), self._w("v." + uid, attn.to_v)
        if cross:
            wcat = self._pack(self.dense, "wqkv." + uid, lambda: torch.cat([wk["w"], wv["w"]], 0).contiguous())
        else:
            wcat = self._pack(self.dense, "wqkv." + uid, lambda: torch.cat([wq["w"], wk["w"], wv["w"]], 0).contiguous())
        if cross:
            uq = self._new(M, C)
            ukv = self._new(Mkv, 2 * C)
            self._mm(("q", uid), xn, C, C, M, wq["w"], C, uq, C, rows_per_sample=hw)
            self._mm(("kv", uid), self.ctx, attn.ctx_dim, attn.ctx_dim, Mkv, wcat, 2 * C, ukv, 2 * C,
                     rows_per_sample=n_kv)
            gq = self._new(M, C)
            gkv = self._new(Mkv, 2 * C)
            K.scale_cols(uq, C, gq, C, B, hw, C, gate, gate.stride(0), 64)
            K.scale_cols(ukv, 2 * C, gkv, 2 * C, B, n_kv, C, gate, gate.stride(0), 64)
            K.scale_cols(ukv[:, C:], 2 * C, gkv[:, C:], 2 * C, B, n_kv, C, gate, gate.stride(0), 64)
            q, ldq, kk, vv, ldkv = gq, C, gkv, gkv[:, C:], 2 * C
            saved = dict(uq=uq, ukv=ukv, gq=gq, gkv=gkv)
        else:
            u = self._new(M, 3 * C)
            self._mm(("qkv", uid), xn, C, C, M, wcat, 3 * C, u, 3 * C, rows_per_sample=hw)
            g = self._new(M, 3 * C)
            for j in range(3):
                K.scale_cols(u[:, j * C:], 3 * C, g[:, j * C:], 3 * C, B, hw, C, gate, gate.stride(0), 64)
            q, ldq, kk, vv, ldkv = g, 3 * C, g[:, C:], g[:, 2 * C:], 3 * C
            saved = dict(u=u, g=g)
        o = self._new(M, C)
        lse = torch.empty(B, heads, hw, device=self.device, dtype=torch.float32)
        sh = self.sched.get(("heads_full", B, heads))
        if sh is None:
            sh = torch.full((B,), heads, device=self.device, dtype=torch.int32)
            self.sched[("heads_full", B, heads)] = sh
        K.attention(q, ldq, kk, ldkv, vv, ldkv, o, C, B, hw, n_kv, sh, heads, 0.125, lse)
        wo = self._w("o." + uid, attn.to_out[0])
        self._mm(("o", uid), o, C, C, M, wo["w"], C, tok_out, C, bias=wo["b"], residual=tok_in, res_ld=C,
                 rows_per_sample=hw)
        self.launches += 8
        self.flops += 4.0 * hw * n_kv * 64 * heads * B
        saved.update(o=o, lse=lse, gate=gate, sh=sh, n_kv=n_kv, cross=cross)
        return saved

    def _b_attn(self, uid: str, attn, gate_idx: int, s: Dict[str, Any], dtok: torch.Tensor, B: int, hw: int, C: int,
                ln_in: Optional[torch.Tensor] = None, ln_g=None, ln_b=None):
        """dtok is d(tok_out); returns d(ln) [M, C] (gradient of the LayerNorm output feeding q / qkv). ln_in / ln_g /
        ln_b: input and affine of that LayerNorm, to recompute its output for the q / k / v weight gradients."""
        M = B * hw
        heads = attn.heads
        n_kv, cross = s["n_kv"], s["cross"]
        Mkv = B * n_kv
        do = self.buf("bw_do", M, C)
        self._mm(("o.T", uid), dtok, C, C, M, self._wT("o." + uid, attn.to_out[0]), C, do, C, rows_per_sample=hw)
        self._wgrad(attn.to_out[0].weight, attn.to_out[0].bias, dtok, C, s["o"], C, M, C, C)
        xn = None
        if self.train_weights:
            xn = self.buf("ln", M, C)
            K.layernorm(ln_in, C, xn, C, M, C, 1e-5, ln_g, ln_b)
        delta = self.buf("bw_delta", B * heads, hw, torch.float32)
        gate = s["gate"]
        gld = gate.stride(0)
        dg = self._dgate(gate_idx)
        if cross:
            dq = self.buf("bw_dq", M, C)
            dkv = self.buf("bw_dkv", Mkv, 2 * C)
            gq, gkv = s["gq"], s["gkv"]
            K.attention_bwd(gq, C, gkv, 2 * C, gkv[:, C:], 2 * C, s["o"], C, do, C, s["lse"], delta, dq, C, dkv, 2 * C,
                            dkv[:, C:], 2 * C, B, hw, n_kv, s["sh"], heads, 0.125)
            duq = self.buf("bw_duq", M, C)
            dukv = self.buf("bw_dukv", Mkv, 2 * C)  # dk, dv only feed the gate gradient (ctx needs none)
            K.scale_cols_bwd(s["uq"], C, dq, C, duq, C, B, hw, C, gate, gld, 64, dg)
            K.scale_cols_bwd(s["ukv"], 2 * C, dkv, 2 * C, dukv, 2 * C, B, n_kv, C, gate, gld, 64, dg)
            K.scale_cols_bwd(s["ukv"][:, C:], 2 * C, dkv[:, C:], 2 * C, dukv[:, C:], 2 * C, B, n_kv, C, gate, gld, 64, dg)
            if self.train_weights:
                self._wgrad(attn.to_q.weight, None, duq, C, xn, C, M, C, C)
                dwkv = torch.zeros(2 * C, attn.ctx_dim, device=self.device, dtype=torch.float32)
                K.wgrad(dukv, 2 * C, self.ctx, attn.ctx_dim, dwkv, None, Mkv, 2 * C, attn.ctx_dim)
                self.wg[id(attn.to_k.weight)], self.wg[id(attn.to_v.weight)] = dwkv[:C], dwkv[C:]
            dln = self.buf("bw_dln", M, C)
            self._mm(("q.T", uid), duq, C, C, M, self._wT("q." + uid, attn.to_q), C, dln, C, rows_per_sample=hw)
        else:
            g = s["g"]
            dqkv = self.buf("bw_dqkv", M, 3 * C)
            K.attention_bwd(g, 3 * C, g[:, C:], 3 * C, g[:, 2 * C:], 3 * C, s["o"], C, do, C, s["lse"], delta, dqkv, 3 * C,
                            dqkv[:, C:], 3 * C, dqkv[:, 2 * C:], 3 * C, B, hw, n_kv, s["sh"], heads, 0.125)
            du = self.buf("bw_du", M, 3 * C)
            for j in range(3):
                K.scale_cols_bwd(s["u"][:, j * C:], 3 * C, dqkv[:, j * C:], 3 * C, du[:, j * C:], 3 * C, B, hw, C,
                                 gate, gld, 64, dg)
            if self.train_weights:
                dwqkv = torch.zeros(3 * C, C, device=self.device, dtype=torch.float32)
                K.wgrad(du, 3 * C, xn, C, dwqkv, None, M, 3 * C, C)
                for j, lin in enumerate((attn.to_q, attn.to_k, attn.to_v)):
                    self.wg[id(lin.weight)] = dwqkv[j * C:(j + 1) * C]
            wcat = self.dense["wqkv." + uid]
            wcat_t = self._pack(self.dense, "wqkv.T." + uid, lambda: wcat.t().contiguous())  # [C, 3C]
            dln = self.buf("bw_dln", M, C)
            self._mm(("qkv.T", uid), du, 3 * C, 3 * C, M, wcat_t, C, dln, C, rows_per_sample=hw)
        self.launches += 8
        return dln

    def t_transformer(self, t: Transformer2DModelWidthGated, x: Act) -> Act:
        B, H, W, hw, C, M = x.B, x.H, x.W, x.hw, x.C, x.rows
        tb = t.transformer_blocks[0]
        cidx = self.gate_cols[t.uid]
        key = ("tr_dense", t.uid)
        dn = self.dense.get(key)
        if dn is None:
            f32 = lambda p: p.detach().to(self.device, torch.float32).contiguous()
            dn = {"g": f32(t.norm.weight), "b": f32(t.norm.bias)}
            for i, ln in enumerate((tb.norm1, tb.norm2, tb.norm3)):
                dn[f"lg{i}"], dn[f"lb{i}"] = f32(ln.weight), f32(ln.bias)
            self.dense[key] = dn
        xn = self.buf("ln", M, C)
        st = self._gn_fwd(x.t, x.ld, C, B, hw, t.groups, 1e-6, dn["g"], dn["b"], xn, False)
        pi = self._w("pi." + t.uid, t.proj_in)
        tok0 = self._new(M, C)
        self._mm(("pi", t.uid), xn, C, C, M, pi["w"], C, tok0, C, bias=pi["b"], rows_per_sample=hw)
        K.layernorm(tok0, C, xn, C, M, C, 1e-5, dn["lg0"], dn["lb0"])
        tok1 = self._new(M, C)
        s1 = self._t_attn(t.uid + ".a1", tb.attn1, cidx["w"][0], xn, tok0, tok1, B, hw, C, cross=False)
        K.layernorm(tok1, C, xn, C, M, C, 1e-5, dn["lg1"], dn["lb1"])
        tok2 = self._new(M, C)
        s2 = self._t_attn(t.uid + ".a2", tb.attn2, cidx["w"][1], xn, tok1, tok2, B, hw, C, cross=True)
        K.layernorm(tok2, C, xn, C, M, C, 1e-5, dn["lg2"], dn["lb2"])
        proj = tb.ff.net[0].proj
        inner = proj.weight.shape[0] // 2
        wp = self._w("ffp." + t.uid, proj)
        hg = self._new(M, 2 * inner)
        self._mm(("ffp", t.uid), xn, C, C, M, wp["w"], 2 * inner, hg, 2 * inner, bias=wp["b"], rows_per_sample=hw)
        fgate = self._soft_gate(cidx["w"][2])
        f = self.buf("ff", M, inner)
        K.geglu(hg, 2 * inner, f, inner, B, hw, inner, fgate, fgate.stride(0), inner // t.gate_width)
        w2 = self._w("ff2." + t.uid, tb.ff.net[2])
        tok3 = self._new(M, C) if self.train_weights else self.buf("tok3", M, C)  # proj_out's input (its wgrad needs it)
        self._mm(("ff2", t.uid), f, inner, inner, M, w2["w"], C, tok3, C, bias=w2["b"], residual=tok2, res_ld=C,
                 rows_per_sample=hw)
        po = self._w("po." + t.uid, t.proj_out)
        y = self._new(M, C)
        self._mm(("po", t.uid), tok3, C, C, M, po["w"], C, y, C, bias=po["b"], residual=x.t, res_ld=x.ld,
                 rows_per_sample=hw)
        out, d = y, None
        if t.depth_gate is not None:
            out = self._new(M, C)
            d = self._soft_depth(cidx["d"])
            K.depth_lerp(x.t, x.ld, y, C, out, C, M, C, d, hw)
        self.launches += 6
        o = Act(out, B, H, W, C)
        self.tape.append(("tr", t, dict(x=x, st=st, tok0=tok0, tok1=tok1, tok2=tok2, s1=s1, s2=s2, hg=hg, fgate=fgate,
                                        inner=inner, y=y, d=d, out=o, dn=dn,
                                        tok3=tok3 if self.train_weights else None)))
        return o

    def b_transformer(self, t: Transformer2DModelWidthGated, s: Dict[str, Any], dout: _G) -> _G:
        x: Act = s["x"]
        B, H, W, hw, C, M = x.B, x.H, x.W, x.hw, x.C, x.rows
        tb = t.transformer_blocks[0]
        cidx = self.gate_cols[t.uid]
        dn = s["dn"]
        if t.depth_gate is not None:
            dy = self._new(M, C)
            dx = self._new(M, C)
            K.depth_lerp_bwd(dout.t, dout.ld, x.t, x.ld, s["y"], C, dy, C, dx, C, False, B, hw, C, s["d"],
                             self.ddepth_T[cidx["d"]])
            K.add_rows(dy, C, dx, C, M, C)  # residual x of proj_out
        else:
            if dout.ld == C:
                dy = dout.t
            else:
                dy = self._new(M, C)
                K.copy_rows(dout.t, dout.ld, dy, C, M, C)
            dx = dy  # out = proj_out(.) + x: the incoming buffer becomes dx after its last read as dy
        dtok = self._new(M, C)
        self._mm(("po.T", t.uid), dy, C, C, M, self._wT("po." + t.uid, t.proj_out), C, dtok, C, rows_per_sample=hw)
        if self.train_weights:
            self._wgrad(t.proj_out.weight, t.proj_out.bias, dy, C, s["tok3"], C, M, C, C)
        # feed-forward
        inner = s["inner"]
        df = self.buf("bw_df", M, inner)
        self._mm(("ff2.T", t.uid), dtok, C, C, M, self._wT("ff2." + t.uid, tb.ff.net[2]), inner, df, inner,
                 rows_per_sample=hw)
        if self.train_weights:  # net.2's input f = GEGLU(hg) is recomputed
            fg = s["fgate"]
            f = self.buf("ff", M, inner)
            K.geglu(s["hg"], 2 * inner, f, inner, B, hw, inner, fg, fg.stride(0), inner // t.gate_width)
            self._wgrad(tb.ff.net[2].weight, tb.ff.net[2].bias, dtok, C, f, inner, M, C, inner)
        dhg = self.buf("bw_dhg", M, 2 * inner)
        K.geglu_bwd(s["hg"], 2 * inner, df, inner, dhg, 2 * inner, B, hw, inner, s["fgate"], s["fgate"].stride(0),
                    inner // t.gate_width, self._dgate(cidx["w"][2]))
        if self.train_weights:
            xn = self.buf("ln", M, C)
            K.layernorm(s["tok2"], C, xn, C, M, C, 1e-5, dn["lg2"], dn["lb2"])
            proj = tb.ff.net[0].proj
            self._wgrad(proj.weight, proj.bias, dhg, 2 * inner, xn, C, M, 2 * inner, C)
        dln = self.buf("bw_dln", M, C)
        self._mm(("ffp.T", t.uid), dhg, 2 * inner, 2 * inner, M, self._wT("ffp." + t.uid, tb.ff.net[0].proj), C, dln, C,
                 rows_per_sample=hw)
        if self.train_weights:
            K.layernorm_affine_bwd(s["tok2"], C, dln, C, M, C, 1e-5, self._affine(tb.norm3))
        K.layernorm_bwd(s["tok2"], C, dln, C, dtok, C, True, M, C, 1e-5, dn["lg2"])
        # cross-attention, self-attention
        dln = self._b_attn(t.uid + ".a2", tb.attn2, cidx["w"][1], s["s2"], dtok, B, hw, C, s["tok1"], dn["lg1"], dn["lb1"])
        if self.train_weights:
            K.layernorm_affine_bwd(s["tok1"], C, dln, C, M, C, 1e-5, self._affine(tb.norm2))
        K.layernorm_bwd(s["tok1"], C, dln, C, dtok, C, True, M, C, 1e-5, dn["lg1"])
        dln = self._b_attn(t.uid + ".a1", tb.attn1, cidx["w"][0], s["s1"], dtok, B, hw, C, s["tok0"], dn["lg0"], dn["lb0"])
        if self.train_weights:
            K.layernorm_affine_bwd(s["tok0"], C, dln, C, M, C, 1e-5, self._affine(tb.norm1))
        K.layernorm_bwd(s["tok0"], C, dln, C, dtok, C, True, M, C, 1e-5, dn["lg0"])
        # proj_in, GroupNorm
        if self.train_weights:
            xn = self.buf("ln", M, C)
            self._gn_again(x.t, x.ld, C, B, hw, t.groups, 1e-6, dn["g"], dn["b"], xn, False, s["st"])
            self._wgrad(t.proj_in.weight, t.proj_in.bias, dtok, C, xn, C, M, C, C)
        dxn = self.buf("bw_dxn", M, C)
        self._mm(("pi.T", t.uid), dtok, C, C, M, self._wT("pi." + t.uid, t.proj_in), C, dxn, C, rows_per_sample=hw)
        self._gn_bwd(x.t, x.ld, C, B, hw, t.groups, 1e-6, s["st"], dn["g"], dn["b"], dxn, dx, C, True, False,
                     daffine=self._affine(t.norm))
        self.launches += 8
        return _G(dx, C, C)

    # ---- samplers ----------------------------------------------------------------------------------
    def t_downsample(self, smp, x: Act) -> Act:
        out = self._new(x.rows // 4, x.C)
        self.conv3x3("ds.%d" % id(smp), smp.conv, x, out, x.C, stride=2)
        o = Act(out, x.B, x.H // 2, x.W // 2, x.C)
        self.tape.append(("down", smp, dict(x=x, out=o)))
        return o

    def b_downsample(self, smp, s, dout: _G) -> _G:
        x: Act = s["x"]
        o: Act = s["out"]
        dyc = dout.t
        if dout.ld != x.C:
            dyc = self._new(o.rows, x.C)
            K.copy_rows(dout.t, dout.ld, dyc, x.C, o.rows, x.C)
        if self.train_weights:
            assert x.ld == x.C, "stride-2 wgrad reads the input through a dense parity view"
            self._wgrad(smp.conv.weight, smp.conv.bias, dyc, x.C, x.t, x.ld, o.rows, x.C, x.C, conv=(x.B, x.H, x.W),
                        stride=2)
        up = self.buf("bw_zi", x.rows, x.C)
        K.zero_insert2x(dyc, up, x.B, o.H, o.W, x.C)
        dx = self._new(x.rows, x.C)
        self._mm(("ds.T", id(smp)), up, x.C, x.C, x.rows, self._wT("ds.%d" % id(smp), smp.conv), x.C, dx, x.C,
                 conv=(x.B, x.H, x.W), k_tap_pitch=x.C, rows_per_sample=x.hw)
        return _G(dx, x.C, x.C)

    def t_upsample(self, smp, x: Act) -> Act:
        up = self.buf("up", x.rows * 4, x.C)
        K.upsample2x(x.t, up, x.B, x.H, x.W, x.C)
        xu = Act(up, x.B, x.H * 2, x.W * 2, x.C)
        out = self._new(xu.rows, x.C)
        self.conv3x3("us.%d" % id(smp), smp.conv, xu, out, x.C)
        o = Act(out, xu.B, xu.H, xu.W, x.C)
        self.tape.append(("up", smp, dict(x=x, out=o)))
        return o

    def b_upsample(self, smp, s, dout: _G) -> _G:
        x: Act = s["x"]
        o: Act = s["out"]
        if self.train_weights:  # the conv's input (nearest-neighbour 2x of x) is recomputed
            xu = self.buf("up", o.rows, x.C)
            K.upsample2x(x.t, xu, x.B, x.H, x.W, x.C)
            self._wgrad(smp.conv.weight, smp.conv.bias, dout.t, dout.ld, xu, x.C, o.rows, x.C, x.C, conv=(o.B, o.H, o.W))
        dup = self.buf("bw_zi", o.rows, x.C)
        self._mm(("us.T", id(smp)), dout.t, dout.ld, x.C, o.rows, self._wT("us.%d" % id(smp), smp.conv), x.C, dup, x.C,
                 conv=(o.B, o.H, o.W), k_tap_pitch=x.C, rows_per_sample=o.hw)
        dx = self._new(x.rows, x.C)
        K.upsample2x_bwd(dup, dx, x.B, x.H, x.W, x.C)
        return _G(dx, x.C, x.C)

    # ------------------------------------------------------------------------------------------
    # backward over the tape
    # ------------------------------------------------------------------------------------------
    def _to_rows(self, g_nchw: torch.Tensor, C: int) -> _G:
        Bn, Cn, Hn, Wn = g_nchw.shape
        t = g_nchw.permute(0, 2, 3, 1).reshape(Bn * Hn * Wn, Cn).to(BF16).contiguous()
        return _G(t, C, C)

    def _acc(self, grads: Dict[int, _G], act: Act, g: Optional[_G]):
        if g is None:
            return
        k = id(act)
        cur = grads.get(k)
        if cur is None:
            grads[k] = g
        else:
            if cur.ld != cur.C:  # accumulate into an owned, dense buffer
                own = self._new(act.rows, cur.C)
                K.copy_rows(cur.t, cur.ld, own, cur.C, act.rows, cur.C)
                cur = _G(own, cur.C, cur.C)
                grads[k] = cur
            K.add_rows(g.t, g.ld, cur.t, cur.ld, act.rows, cur.C)

    def backward(self, dpred: Optional[torch.Tensor], dtaps: List[Optional[torch.Tensor]],
                 tape_id: Optional[int] = None) -> torch.Tensor:
        """Returns d(loss)/d(arch) [gate_rows, 1620] in fp32."""
        if tape_id is not None and tape_id != self.tape_id:
            raise RuntimeError(
                f"backward of U-Net forward #{tape_id} requested, but the activation tape now holds forward #{self.tape_id}: "
                "the training engine keeps ONE taped forward per model (run backward before the next grad-enabled "
                "forward, as the reference's train loops do: trainer.py:918-933)")
        m = self.m
        grads: Dict[int, _G] = {}
        x, stats, dn = self.final
        B = self.B
        groups = m.config["norm_num_groups"]
        if dpred is not None:
            cout = m.config["out_channels"]
            dp = torch.zeros(x.rows, 64, device=self.device, dtype=BF16)
            dp[:, :cout] = dpred.permute(0, 2, 3, 1).reshape(x.rows, cout).to(BF16)
            da = self.buf("bw_a", x.rows, x.C)
            self._mm(("conv_out.T",), dp, 64, 64, x.rows, self._wT("conv_out", m.conv_out, pad_out_to=64), x.C, da, x.C,
                     conv=(x.B, x.H, x.W), k_tap_pitch=64, rows_per_sample=x.hw)
            if self.train_weights:  # conv_out: input silu(GN(x)) recomputed; dy is the 64-column padded prediction gradient
                a = self.buf("gn_a", x.rows, x.C)
                self._gn_again(x.t, x.ld, x.C, B, x.hw, groups, m.config["norm_eps"], dn["g"], dn["b"], a, True, stats)
                dw64 = torch.zeros(64, 9 * x.C, device=self.device, dtype=torch.float32)
                db64 = torch.zeros(64, device=self.device, dtype=torch.float32)
                K.wgrad(dp, 64, a, x.C, dw64, db64, x.rows, 64, x.C, conv=(x.B, x.H, x.W))
                self.wg[id(m.conv_out.weight)], self.wg[id(m.conv_out.bias)] = dw64[:cout], db64[:cout]
            dxl = self._new(x.rows, x.C)
            self._gn_bwd(x.t, x.ld, x.C, B, x.hw, groups, m.config["norm_eps"], stats, dn["g"], dn["b"], da, dxl, x.C,
                         False, True, daffine=self._affine(m.conv_norm_out))
            self._acc(grads, x, _G(dxl, x.C, x.C))
        for act, dt in zip(self.tap_acts, dtaps):
            if dt is not None:
                self._acc(grads, act, self._to_rows(dt, act.C))
        for kind, mod, s in reversed(self.tape):
            out: Act = s["out"]
            g = grads.pop(id(out), None)
            if g is None:
                continue
            if kind == "res":
                dxin = self.b_resnet(mod, s, g)
                if dxin is None:
                    continue
                if s["skip"] is not None:
                    xC = s["x"].C
                    self._acc(grads, s["x"], _G(dxin.t, dxin.ld, xC))
                    self._acc(grads, s["skip"], _G(dxin.t[:, xC:], dxin.ld, s["skip"].C))
                else:
                    self._acc(grads, s["x"], dxin)
            elif kind == "tr":
                self._acc(grads, s["x"], self.b_transformer(mod, s, g))
            elif kind == "down":
                self._acc(grads, s["x"], self.b_downsample(mod, s, g))
            elif kind == "up":
                self._acc(grads, s["x"], self.b_upsample(mod, s, g))
        if self.train_weights:
            self._finish_weight_grads(grads)
        self.darch[:, self.n_width:].add_(self.ddepth_T.t())
        d = self.darch
        if self.gate_rows != B:  # gates were tiled along the batch (gates.py:18-19): fold the copies back
            d = d.reshape(B // self.gate_rows, self.gate_rows, -1).sum(0)
        self.tape = []
        self.final = None
        return d


def _silu_grad(z: torch.Tensor) -> torch.Tensor:
    sg = torch.sigmoid(z)
    return sg * (1.0 + z * (1.0 - sg))


def _finish_weight_grads(self: TrainEngine, grads: Dict[int, _G]) -> None:
    """conv_in, the time-embedding chain, and the conversion of the kernel-layout accumulators to parameter shapes."""
    m = self.m
    B = self.B
    c0 = m.config["block_out_channels"][0]
    cin = m.config["in_channels"]
    tdim = c0 * 4
    # conv_in: linear wgrad over the im2col rows ([rows, 64], OHWI tap order, K = 9 * cin zero-padded to 64)
    g0 = grads.pop(id(self.x0_act), None)
    if g0 is not None:
        dw_in = torch.zeros(c0, 64, device=self.device, dtype=torch.float32)
        db_in = torch.zeros(c0, device=self.device, dtype=torch.float32)
        K.wgrad(g0.t, g0.ld, self.im2col, 64, dw_in, db_in, self.x0_act.rows, c0, 64)
        self.wg[id(m.conv_in.weight)], self.wg[id(m.conv_in.bias)] = dw_in[:, :9 * cin], db_in
    # time embedding: h1 += (time_emb_proj(silu(temb)) + time_emb_proj.bias + conv1.bias)[b] in all 22 ResNets
    dproj = self.dproj_all                                   # [B, temb_total] fp32 (per-sample column sums of dh1)
    ntot = m._temb_total
    dproj16 = dproj.to(BF16)
    te = self.buf("t_e", B, tdim)                            # silu(temb), bf16
    dw_all = torch.zeros(ntot, tdim, device=self.device, dtype=torch.float32)
    K.wgrad(dproj16, ntot, te, tdim, dw_all, None, B, ntot, tdim)
    db_all = dproj.sum(0)
    for r in m._resnets:
        off = m._temb_off[r.uid]
        self.wg[id(r.time_emb_proj.weight)] = dw_all[off:off + r.cout]
        self.wg[id(r.time_emb_proj.bias)] = db_all[off:off + r.cout]
        self.wg[id(r.conv1.bias)] = db_all[off:off + r.cout]
    wt_all = self._pack(self.dense, "temb_all.T", lambda: self._temb_pack()["w"][:ntot].t().contiguous())  # [tdim, temb_total]
    dte = self._new(B, tdim)
    self._mm(("temb.T",), dproj16, ntot, ntot, B, wt_all, tdim, dte, tdim)
    te_m = m.time_embedding
    dz2 = (dte.float() * _silu_grad(self.te_z2)).to(BF16)
    self._wgrad(te_m.linear_2.weight, te_m.linear_2.bias, dz2, tdim, self.buf("t_h", B, tdim), tdim, B, tdim, tdim)
    dh = self._new(B, tdim)
    self._mm(("te2.T",), dz2, tdim, tdim, B, self._wT("te2", te_m.linear_2), tdim, dh, tdim)
    dz1 = (dh.float() * _silu_grad(self.te_z1)).to(BF16)
    self._wgrad(te_m.linear_1.weight, te_m.linear_1.bias, dz1, tdim, self.buf("t_sin", B, c0), c0, B, tdim, c0)
    # kernel layout -> parameter layout
    out: Dict[int, torch.Tensor] = {}
    pruned_beta = {}
    if m.pruned_semantics:  # norm2.bias of a removed channel does not exist in the sliced model: no gradient
        pruned_beta = {id(r.norm2): self._resnet_keep_mask(r).reshape(-1) for r in m._resnets}
    for mod in m.modules():
        if isinstance(mod, (nn.GroupNorm, nn.LayerNorm)):
            g = self.wg.get(id(mod.weight))
            if g is not None:
                out[id(mod.weight)], out[id(mod.bias)] = g[:, 0], g[:, 1]
                if id(mod) in pruned_beta:
                    out[id(mod.bias)] = g[:, 1] * pruned_beta[id(mod)]
        elif isinstance(mod, nn.Conv2d):
            g = self.wg.get(id(mod.weight))
            if g is not None:
                co, ci, kh, kw = mod.weight.shape
                out[id(mod.weight)] = g.reshape(co, kh, kw, ci).permute(0, 3, 1, 2)   # OHWI -> OIHW
            if mod.bias is not None and id(mod.bias) in self.wg:
                out[id(mod.bias)] = self.wg[id(mod.bias)]
        elif isinstance(mod, nn.Linear):
            if id(mod.weight) in self.wg:
                out[id(mod.weight)] = self.wg[id(mod.weight)]
            if mod.bias is not None and id(mod.bias) in self.wg:
                out[id(mod.bias)] = self.wg[id(mod.bias)]
    self.param_grads = out


TrainEngine._finish_weight_grads = _finish_weight_grads


class UNetFineTuneFunction(torch.autograd.Function):
    """(prediction, 9 block activations) = U-Net(sample, t, ctx; weights) with gradients to EVERY U-Net parameter: the
    student forward / backward of the fine-tune stage (trainer.py:1727-1729). The gates are constants here (one static
    expert or the all-ones teacher layout); weights are re-packed from the parameters on every call."""

    @staticmethod
    def forward(ctx, model, sample, timestep, enc, n_taps_out, *params):
        eng = model._get_train_engine(sample.device)
        eng.train_weights = True
        model._engine = None  # the inference engine's packed weights go stale as soon as the optimizer steps
        # the parameters changed since the last step (optimizer): re-derive the packed bf16 copies -- in place, as one
        # CUDA-graph replay of the registered builders (first call: nothing is packed yet, the forward builds them)
        if getattr(eng, "_ft_packs_ready", False):
            eng.refresh_packs()
        else:
            eng.dense, eng.expert, eng._packs, eng._pack_graph = {}, {}, [], None
            eng._ft_packs_ready = True
        flat_w, flat_d = model._flat_gates
        y, taps = eng.run_train(sample, timestep, enc, [g.detach() for g in list(flat_w) + list(flat_d)])
        ctx.eng = eng
        ctx.tape_id = eng.tape_id
        ctx.params = params
        return tuple([y] + [t.nchw() for t in taps])

    @staticmethod
    def backward(ctx, dy, *dtaps):
        eng = ctx.eng
        eng.backward(dy, list(dtaps), tape_id=ctx.tape_id)
        eng.train_weights = False
        pg = eng.param_grads
        grads = []
        for p in ctx.params:
            g = pg.get(id(p))
            grads.append(None if g is None else g.to(p.dtype).reshape(p.shape).contiguous())
        return (None, None, None, None, None, *grads)


class UNetTrainFunction(torch.autograd.Function):
    """(prediction, 9 block activations) = U-Net(sample, t, ctx; gates) with gradients to the gates only."""

    @staticmethod
    def forward(ctx, model, sample, timestep, enc, n_taps_out, *gates):
        eng = model._get_train_engine(sample.device)
        y, taps = eng.run_train(sample, timestep, enc, list(gates))
        ctx.eng = eng
        ctx.tape_id = eng.tape_id
        ctx.gate_shapes = [g.shape for g in gates]
        ctx.n_width_gates = len(eng.width_starts) - 1
        outs = [y] + [t.nchw() for t in taps]  # bf16 channels-last views of the taped block outputs (zero-copy)
        return tuple(outs)

    @staticmethod
    def backward(ctx, dy, *dtaps):
        eng = ctx.eng
        darch = eng.backward(dy, list(dtaps), tape_id=ctx.tape_id)
        grads = []
        ws = eng.width_starts
        for i, shp in enumerate(ctx.gate_shapes):
            if i < ctx.n_width_gates:
                g = darch[:, ws[i]:ws[i + 1]]
            else:
                g = darch[:, eng.n_width + (i - ctx.n_width_gates)]
            grads.append(g.reshape(shp).contiguous())
        return (None, None, None, None, None, *grads)
