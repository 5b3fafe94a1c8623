#!/usr/bin/env python
"""CUDA-event timing of aptp_attention_bwd (delta + dQ + dK/dV kernels) at the train step's shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_pruning_b200 import kernels as K

DEV = "cuda"


def run(B, heads, Nq, Nkv, iters=10):
    C = heads * 64
    g = torch.Generator(device=DEV).manual_seed(0)
    q, do = (torch.randn(B * Nq, C, device=DEV, generator=g).bfloat16() for _ in range(2))
    k, v = (torch.randn(B * Nkv, C, device=DEV, generator=g).bfloat16() for _ in range(2))
    out = torch.zeros_like(q)
    sh = torch.full((B,), heads, device=DEV, dtype=torch.int32)
    lse = torch.zeros(B, heads, Nq, device=DEV)
    K.attention(q, C, k, C, v, C, out, C, B, Nq, Nkv, sh, heads, 0.125, lse)
    delta = torch.empty(B, heads, Nq, device=DEV)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    f = lambda: K.attention_bwd(q, C, k, C, v, C, out, C, do, C, lse, delta, dq, C, dk, C, dv, C, B, Nq, Nkv, sh, heads, 0.125)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    K.check_abort()
    ms = e0.elapsed_time(e1) / iters
    flop = 7 * 2.0 * B * heads * Nq * Nkv * 64  # 3 (dQ kernel) + 4 (dK/dV kernel) matmuls
    print(f"attn_bwd B{B} h{heads} Nq{Nq} Nkv{Nkv}: {ms:.3f} ms  {flop / ms / 1e9:.0f} TFLOP/s (7 matmuls)")


if __name__ == "__main__":
    run(32, 5, 4096, 4096)
    run(32, 5, 4096, 77)
    run(32, 10, 1024, 1024)
    run(32, 10, 1024, 77)
    run(32, 20, 256, 256)
