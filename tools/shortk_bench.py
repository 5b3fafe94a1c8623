"""Micro-benchmark of the level-0 (64x64, C=320) transformer GEMMs with the epilogues the engine uses, against their HBM
floor (CUDA events; run on the GPU box). One line per shape:  python tools/shortk_bench.py [names...]
Kernel-tuning builds are selected with APTP_LIB=variants/libaptp_<name>.so (tools/build_variant.sh)."""
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import torch  # noqa: E402

from diffusion_pruning_b200 import kernels as K  # noqa: E402
from diffusion_pruning_b200 import _lib  # noqa: E402

M = int(os.environ.get("SHORTK_M", 262144))
HW = 4096
PEAK = 6540.0  # GB/s, MEASURED_PEAKS.json burst copy figure


def timeit(fn, iters=20):
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


def run(name, Kd, N, bn, *, geglu=False, ln=False, res=None, out_f32=False, rowstat=False, colstat=False, nseg=1):
    dev = "cuda"
    a = torch.randn(M, Kd, device=dev).bfloat16()
    w = (torch.randn(N, Kd, device=dev) * 0.05).bfloat16()
    n_out = N // 2 if geglu else N
    out = torch.empty(M, n_out, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    bias = torch.randn(N, device=dev)
    residual = None
    if res == "bf16":
        residual = torch.randn(M, n_out, device=dev).bfloat16()
    elif res == "f32":
        residual = torch.randn(M, n_out, device=dev)
    flags = (_lib.EPI_GEGLU if geglu else 0) | (_lib.EPI_RES_F32 if res == "f32" else 0)
    kw = {}
    if ln:
        kw["ln_colsum"] = torch.randn(N, device=dev)
        kw["ln_rowstats"] = torch.rand(M, 2, device=dev) + 0.5
    if rowstat:
        kw["rowstat_out"] = torch.empty(M, n_out // 32, 2, device=dev)
    if colstat:
        kw["colstat"] = (torch.empty(M // 32, n_out, device=dev), torch.empty(M // 32, n_out, device=dev))
    rows = M // nseg
    segs = [K.Segment(i * rows, (i + 1) * rows, n_out, Kd // 64) for i in range(nseg)]
    sched = K.build_schedule(segs, bn, dev, geglu=geglu)
    ms = timeit(lambda: K.grouped_gemm(a, w, out, sched, a_ld=Kd, a_k=Kd, a_rows=M, out_ld=n_out, bias=bias,
                                       residual=residual, res_ld=n_out, flags=flags, rows_per_sample=HW,
                                       out_mode=_lib.OUT_F32 if out_f32 else _lib.OUT_BF16, **kw))
    fl = 2.0 * M * Kd * N
    by = M * Kd * 2 + N * Kd * 2 + out.numel() * out.element_size()
    if residual is not None:
        by += residual.numel() * residual.element_size()
    if rowstat:
        by += kw["rowstat_out"].numel() * 4
    if ln:
        by += M * 8
    floor = by / PEAK / 1e6
    print(f"{name:8s} K{Kd} N{N} bn{bn} tiles{sched.n_tiles}: {ms*1e3:7.1f} us  {fl/ms/1e9:6.0f} TFLOP/s  {by/ms/1e6:5.0f} GB/s  "
          f"floor {floor*1e3:6.1f} us ({floor/ms:.2f})", flush=True)


SHAPES = {
    "pi": lambda: run("pi", 320, 320, 160, rowstat=True),
    "plain": lambda: run("plain", 320, 320, 160),
    "qkv": lambda: run("qkv", 320, 960, 128, ln=True),
    "to_out": lambda: run("to_out", 320, 320, 160, res="bf16", rowstat=True),
    "geglu": lambda: run("geglu", 320, 2560, 256, geglu=True, ln=True),
    "geglu192": lambda: run("geglu192", 320, 2304, 192, geglu=True, ln=True),
    "ffout": lambda: run("ffout", 1280, 320, 160, res="bf16"),
    "po": lambda: run("po", 320, 320, 160, res="f32", out_f32=True, colstat=True),
    "po_nostat": lambda: run("po_nostat", 320, 320, 160, res="f32", out_f32=True),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(SHAPES)
    for n in names:
        SHAPES[n]()
    K.check_abort()
