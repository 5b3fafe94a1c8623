"""Micro-benchmark of the HBM-bound normalisation kernels (CUDA events; run on the GPU box)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402
from diffusion_pruning_b200 import kernels as K  # noqa: E402
from gemm_bench import timeit  # noqa: E402


def bench_ln(rows, C):
    x = torch.randn(rows, C, device="cuda").bfloat16()
    y = torch.empty_like(x)
    g = torch.ones(C, device="cuda")
    b = torch.zeros(C, device="cuda")
    ms = timeit(lambda: K.layernorm(x, C, y, C, rows, C, 1e-5, g, b), iters=20)
    print(f"layernorm rows{rows} C{C}: {ms*1e3:.1f} us {2*rows*C*2/ms/1e6:.0f} GB/s", flush=True)


def bench_gn(B, hw, C, silu):
    x = torch.randn(B * hw, C, device="cuda").bfloat16()
    y = torch.empty_like(x)
    g = torch.ones(C, device="cuda")
    b = torch.zeros(C, device="cuda")
    stats = torch.zeros(B, 32, 2, device="cuda")
    gs = C // 32

    def st():
        stats.zero_()
        K.groupnorm_stats(x, C, C, None, 0, 0, B, hw, gs, None, stats, 32)
    ms_s = timeit(st, iters=20)
    ms_a = timeit(lambda: K.groupnorm_apply(x, C, C, None, 0, 0, y, C, B, hw, gs, 1e-5, stats, 32, g, b, C, None, None,
                                            None, 32, silu), iters=20)
    n = B * hw * C * 2
    print(f"groupnorm B{B} hw{hw} C{C} silu={int(silu)}: stats {ms_s*1e3:.1f} us {n/ms_s/1e6:.0f} GB/s | "
          f"apply {ms_a*1e3:.1f} us {2*n/ms_a/1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    bench_ln(262144, 320)
    bench_ln(65536, 640)
    bench_ln(16384, 1280)
    bench_gn(64, 4096, 320, True)
    bench_gn(64, 4096, 320, False)
    bench_gn(64, 4096, 960, True)
    bench_gn(64, 1024, 640, True)
    bench_gn(64, 256, 1280, True)
    bench_gn(64, 64, 2560, True)
