"""Micro-benchmark of the tcgen05 grouped GEMM / conv kernel (CUDA events; run on the GPU box)."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch  # noqa: E402

from diffusion_pruning_b200 import kernels as K  # noqa: E402
from diffusion_pruning_b200._lib import A_CONV3X3  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def bench_linear(M, Kd, N, bn, residual=False, geglu=False):
    from diffusion_pruning_b200._lib import EPI_GEGLU
    a = torch.randn(M, Kd, device="cuda").bfloat16()
    w = torch.randn(N, Kd, device="cuda").bfloat16()
    n_out = N // 2 if geglu else N
    out = torch.empty(M, n_out, device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, n_out, device="cuda").bfloat16() if residual else None
    sched = K.build_schedule([K.Segment(0, M, n_out, Kd // 64)], bn, "cuda", geglu=geglu)
    ms = timeit(lambda: K.grouped_gemm(a, w, out, sched, a_ld=Kd, a_k=Kd, a_rows=M, out_ld=n_out, bias=bias,
                                       residual=res, res_ld=n_out, flags=EPI_GEGLU if geglu else 0))
    ms_t = timeit(lambda: torch.matmul(a, w.t()))
    fl = 2.0 * M * Kd * N
    gb = (M * Kd + M * n_out * (2 if residual else 1) + N * Kd) * 2 / 1e9
    print(f"linear M{M} K{Kd} N{N} bn{bn} res={int(residual)} geglu={int(geglu)}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s "
          f"{gb/ms*1e3:.0f} GB/s | torch.matmul {ms_t:.3f} ms {fl/ms_t/1e9:.0f} TFLOP/s", flush=True)


def bench_conv(B, H, C, Cout, bn):
    x = torch.randn(B * H * H, C, device="cuda").bfloat16()
    w = torch.randn(Cout, 9 * C, device="cuda").bfloat16()
    out = torch.empty(B * H * H, Cout, device="cuda", dtype=torch.bfloat16)
    sched = K.build_schedule([K.Segment(0, B * H * H, Cout, C // 64)], bn, "cuda", mode=A_CONV3X3, Ho=H, Wo=H)
    ms = timeit(lambda: K.grouped_gemm(x, w, out, sched, a_ld=C, a_k=C, a_rows=B * H * H, mode=A_CONV3X3, batch=B, H=H,
                                       W=H, k_tap_pitch=C, out_ld=Cout, rows_per_sample=H * H))
    fl = 2.0 * B * H * H * 9 * C * Cout
    print(f"conv3x3 B{B} {H}x{H} {C}->{Cout} bn{bn}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s", flush=True)


def bench_attn(B, heads, N, Nkv):
    C = heads * 64
    q = torch.randn(B * N, C, device="cuda").bfloat16()
    k = torch.randn(B * Nkv, C, device="cuda").bfloat16()
    v = torch.randn(B * Nkv, C, device="cuda").bfloat16()
    o = torch.empty(B * N, C, device="cuda", dtype=torch.bfloat16)
    sh = torch.full((B,), heads, device="cuda", dtype=torch.int32)
    ms = timeit(lambda: K.attention(q, C, k, C, v, C, o, C, B, N, Nkv, sh, heads, 0.125))
    fl = 4.0 * B * heads * N * Nkv * 64
    print(f"attention B{B} h{heads} N{N} Nkv{Nkv}: {ms:.3f} ms {fl/ms/1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    bench_linear(8192, 8192, 8192, 256)
    bench_linear(262144, 320, 320, 160)
    bench_linear(262144, 320, 320, 160, residual=True)
    bench_linear(262144, 320, 2560, 256)
    bench_linear(262144, 320, 2560, 256, geglu=True)
    bench_linear(262144, 320, 2560, 192, geglu=True)
    bench_linear(262144, 1280, 320, 160, residual=True)
    bench_linear(65536, 640, 640, 160)
    bench_linear(16384, 1280, 1280, 256)
    bench_conv(64, 64, 320, 320, 160)
    bench_conv(64, 32, 640, 640, 160)
    bench_conv(64, 16, 1280, 1280, 256)
    bench_conv(64, 8, 1280, 1280, 256)
    bench_attn(64, 5, 4096, 4096)
    bench_attn(64, 10, 1024, 1024)
    bench_attn(64, 5, 4096, 77)
    K.check_abort()
