#!/usr/bin/env python
"""CUDA-event timing of aptp_wgrad at fine-tune shapes (batch 32, SD-2.1 levels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusion_pruning_b200 import kernels as K

DEV = "cuda"


def run(name, rows, n_out, k_in, conv=None, iters=5, splits=0):
    g = torch.Generator(device=DEV).manual_seed(0)
    dy = torch.randn(rows, n_out, device=DEV, generator=g).bfloat16()
    a = torch.randn(rows, k_in, device=DEV, generator=g).bfloat16()
    taps = 9 if conv else 1
    dw = torch.zeros(n_out, taps * k_in, device=DEV)
    db = torch.zeros(n_out, device=DEV)
    f = lambda: K.wgrad(dy, n_out, a, k_in, dw, db, rows, n_out, k_in, conv=conv, splits=splits)
    for _ in range(2):
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
    flop = 2.0 * rows * n_out * k_in * taps
    print(f"wgrad {name:28s} rows{rows} n{n_out} k{k_in} taps{taps} splits{splits}: {ms:.3f} ms  {flop / ms / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    run("conv L0 320->320", 32 * 64 * 64, 320, 320, conv=(32, 64, 64))
    run("conv L1 640->640", 32 * 32 * 32, 640, 640, conv=(32, 32, 32))
    run("conv L2 1280->1280", 32 * 16 * 16, 1280, 1280, conv=(32, 16, 16))
    run("linear ff.out L0", 32 * 64 * 64, 320, 1280)
    run("linear ff.proj L0", 32 * 64 * 64, 2560, 320)
    run("linear qkv L1", 32 * 32 * 32, 1920, 640)
