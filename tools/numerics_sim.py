"""CPU experiment (no GPU): which bf16 storage points of the engine dominate the whole-U-Net error?

Re-runs the fp32 oracle with bf16 rounding emulated at the places where the CUDA engine stores bf16
(weights, GEMM A operands, GEMM outputs, the residual stream, q/k/v, P, attention output, GEGLU output) and
prints cosine / max-abs against the un-rounded fp32 oracle for several knob settings. Used to decide which
tensors to keep in fp32. Run: python tools/numerics_sim.py [--full] [--H 32]
"""
import argparse
import copy
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import unet_oracle as O  # noqa: E402
from diffusion_pruning_b200.synthetic import split_arch, synthetic_codes  # noqa: E402


class Knobs:
    w = True        # weights bf16
    act = True      # GEMM A operands (norm outputs) bf16
    stream = True   # residual stream tensors bf16
    mid = True      # conv1 output h1 bf16
    qkv = True
    p = True        # softmax probabilities bf16
    o = True        # attention output bf16
    ff = True       # GEGLU output bf16
    sc = True       # shortcut conv output bf16 (before being added in conv2's epilogue)
    tok_f32 = False  # transformer token stream (tok) fp32 only inside the transformer
    tok_bf16 = False  # force the token stream to bf16 even when stream is fp32 (block-level fp32 stream only)


KN = Knobs()
KN.fp32_w = set()


def rs_(x):
    """block-level stream rounding, skipped at the spatial widths listed in KN.fp32_w"""
    if x.shape[-1] in KN.fp32_w:
        return x
    return r(x, KN.stream)


def r(x, on=True):
    return x.bfloat16().float() if on else x


def resnet_fwd(self, x, temb):
    x_in = x[:, : x.shape[1] - self.skip_dim] if (self.depth_gated and self.skip_dim) else x
    if self.dropped:
        return x_in
    h = r(F.silu(self.norm1(x)), KN.act)
    h = self.conv1(h)
    h = h + self.time_emb_proj(r(F.silu(temb), KN.act))[:, :, None, None]
    h = r(h, KN.mid)
    if not self.pruned:
        h = O.width_gate(h, self.gate)
    h = r(F.silu(self.norm2(h)), KN.act)
    h = self.conv2(h)
    sc = r(self.conv_shortcut(x), KN.sc) if self.conv_shortcut is not None else x
    out = rs_(sc + h)
    if self.depth_gated and not self.pruned:
        out = rs_(O.depth_gate(x_in, out, self.depth))
    return out


def attn_fwd(self, x, ctx=None):
    B = x.shape[0]
    ctx = x if ctx is None else ctx
    hd = self.dim // self.heads
    q = r(self.to_q(x), KN.qkv).view(B, -1, self.heads, hd).transpose(1, 2)
    k = r(self.to_k(ctx), KN.qkv).view(B, -1, self.heads, hd).transpose(1, 2)
    v = r(self.to_v(ctx), KN.qkv).view(B, -1, self.heads, hd).transpose(1, 2)
    q, k, v = O.width_gate(q, self.gate), O.width_gate(k, self.gate), O.width_gate(v, self.gate)
    s = (q @ k.transpose(-1, -2)) / 8.0
    m = s.max(-1, keepdim=True).values
    p = torch.exp(s - m)
    l = p.sum(-1, keepdim=True)
    o = (r(p, KN.p) @ v) / l
    o = r(o, KN.o).transpose(1, 2).reshape(B, -1, self.dim)
    return self.to_out[0](o)


def ff_fwd(self, x):
    h, g = self.net[0].proj(x).chunk(2, dim=-1)
    h, g = O.linear_width_gate(h, self.gate), O.linear_width_gate(g, self.gate)
    return self.net[2](r(h * F.gelu(g), KN.ff))


def block_fwd(self, x, ctx):
    rs = (KN.stream and not KN.tok_f32) or KN.tok_bf16
    x = r(self.attn1(r(self.norm1(x), KN.act)) + x, rs)
    x = r(self.attn2(r(self.norm2(x), KN.act), ctx) + x, rs)
    x = r(self.ff(r(self.norm3(x), KN.act)) + x, rs)
    return x


def transformer_fwd(self, x, ctx):
    B, C, H, W = x.shape
    h = r(self.norm(x), KN.act)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    h = r(self.proj_in(h), (KN.stream and not KN.tok_f32) or KN.tok_bf16)
    for blk in self.transformer_blocks:
        h = blk(h, r(ctx, KN.act))
    h = self.proj_out(r(h, KN.act))  # proj_out reads tok as a bf16 A operand
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    out = rs_(h + x)
    if self.depth_gated:
        out = rs_(O.depth_gate(x, out, self.depth))
    return out


def unet_fwd(self, sample, timestep, ctx):
    t = timestep.expand(sample.shape[0])
    te = self.time_embedding
    emb = r(O.timestep_sinusoid(t, self.cfg.block_out_channels[0]), KN.act)
    temb = te.linear_2(r(F.silu(te.linear_1(emb)), KN.act))
    x = rs_(self.conv_in(r(sample, KN.act)))
    skips = [x]
    for blk in self.down_blocks:
        outs = ()
        for i, rs in enumerate(blk.resnets):
            x = rs(x, temb)
            if blk.attentions is not None:
                x = blk.attentions[i](x, ctx)
            outs += (x,)
        if blk.downsamplers is not None:
            x = rs_(blk.downsamplers[0].conv(r(x, KN.act)))
            outs += (x,)
        skips += list(outs)
    x = self.mid_block(x, temb, ctx)
    for blk in self.up_blocks:
        for i, rs in enumerate(blk.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = rs(x, temb)
            if blk.attentions is not None:
                x = blk.attentions[i](x, ctx)
        if blk.upsamplers is not None:
            x = rs_(blk.upsamplers[0].conv(r(F.interpolate(x, scale_factor=2.0, mode="nearest"), KN.act)))
    return self.conv_out(r(F.silu(self.conv_norm_out(x)), KN.act))


def patch():
    O.Resnet.forward = resnet_fwd
    O.Attention.forward = attn_fwd
    O.FeedForward.forward = ff_fwd
    O.BasicTransformerBlock.forward = block_fwd
    O.Transformer.forward = transformer_fwd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--H", type=int, default=32)
    ap.add_argument("--B", type=int, default=4)
    ap.add_argument("--beta", type=float, default=0.0)
    ap.add_argument("--codes", type=str, default="0,3,3,7")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    cfg = O.UNetConfig() if a.full else O.UNetConfig.tiny()
    oracle = O.GatedUNetOracle(cfg).eval()
    O.seeded_init(oracle, 0, a.beta)
    st = oracle.get_structure()
    codes = synthetic_codes(st, 8)
    if a.codes == "ones":
        arch = torch.ones(a.B, codes.shape[1])
    else:
        ids = [int(c) for c in a.codes.split(",")][: a.B]
        arch = codes[ids]
    g = torch.Generator().manual_seed(1)
    sample = torch.randn(a.B, 4, a.H, a.H, generator=g)
    ctx = torch.randn(a.B, 77, cfg.cross_attention_dim, generator=g)
    t = torch.tensor([981, 661, 341, 21] * ((a.B + 3) // 4))[: a.B]
    oracle.set_structure(split_arch(arch.clone(), st))
    with torch.no_grad():
        ref = oracle(sample, t, ctx)
    print(f"ref: std={ref.std():.4f} absmax={ref.abs().max():.4f}")
    fwd_ref = (O.Resnet.forward, O.Attention.forward, O.FeedForward.forward, O.BasicTransformerBlock.forward,
               O.Transformer.forward)
    patch()
    q = copy.deepcopy(oracle)
    experiments = [
        ("all bf16 (engine today)", {}),
        ("weights fp32", {"w": False}),
        ("stream fp32", {"stream": False}),
        ("tok fp32 (inside transformer only)", {"tok_f32": True}),
        ("block-level stream fp32, tok bf16", {"stream": False, "tok_bf16": True}),
        ("block-level stream + sc fp32, tok bf16", {"stream": False, "tok_bf16": True, "sc": False}),
        ("block-level fp32 at level 0 only", {"fp32_w": "0"}),
        ("block-level fp32 at levels 1-3 only", {"fp32_w": "123"}),
        ("block-level fp32 at levels 1-3 + sc", {"fp32_w": "123", "sc": False}),
        ("stream+mid fp32", {"stream": False, "mid": False}),
        ("stream+mid+sc fp32", {"stream": False, "mid": False, "sc": False}),
        ("stream+mid+sc+qkv+o+ff fp32", {"stream": False, "mid": False, "sc": False, "qkv": False, "o": False, "ff": False}),
        ("only weights bf16", {"stream": False, "mid": False, "sc": False, "qkv": False, "o": False, "ff": False,
                               "act": False, "p": False}),
        ("only act bf16", {"w": False, "stream": False, "mid": False, "sc": False, "qkv": False, "o": False,
                           "ff": False, "p": False}),
        ("nothing", {"w": False, "stream": False, "mid": False, "sc": False, "qkv": False, "o": False, "ff": False,
                     "act": False, "p": False}),
    ]
    for name, kn in experiments:
        for k in ("w", "act", "stream", "mid", "qkv", "p", "o", "ff", "sc"):
            setattr(KN, k, True)
        KN.tok_f32 = False
        KN.tok_bf16 = False
        KN.fp32_w = set()
        for k, v in kn.items():
            if k == "fp32_w":
                v = {a.H >> int(c) for c in v}
            setattr(KN, k, v)
        m = copy.deepcopy(oracle)
        if KN.w:
            with torch.no_grad():
                for n_, p_ in m.named_parameters():
                    if p_.ndim >= 2:
                        p_.copy_(p_.bfloat16().float())
        m.set_structure(split_arch(arch.clone(), st))
        with torch.no_grad():
            got = unet_fwd(m, sample, t, ctx)
        d = (got - ref).abs()
        cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
        rel = ((got - ref).norm() / ref.norm()).item()
        print(f"{name:44s} cos={cos:.6f} 1-cos={1 - cos:.2e} rel_l2={rel:.4f} max_abs={d.max():.4f} "
              f"max_abs/scale={d.max() / max(1.0, ref.abs().max()):.4f}")
    del q, fwd_ref


if __name__ == "__main__":
    main()
