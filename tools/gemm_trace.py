"""Role-level stall accounting of the grouped GEMM (kernel-tuning build with -DAPTP_GEMM_TRACE):
  APTP_LIB=variants/libaptp_trace.so python tools/gemm_trace.py pi qkv geglu po
Per shape: average cycles per CTA each role spent in total and blocked on each barrier, per tile."""
import ctypes as C
import os
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import shortk_bench as S  # noqa: E402
from diffusion_pruning_b200 import _lib  # noqa: E402

lib = _lib.load()
fn = lib.aptp_debug_gemm_trace
fn.argtypes = [C.c_void_p, C.c_int]
fn.restype = C.c_int
S.timeit = lambda f, iters=1: (f(), torch.cuda.synchronize(), 1.0)[2]   # one launch per shape


def dump(name):
    buf = np.zeros((512, 16), dtype=np.int64)
    torch.cuda.synchronize()
    assert fn(buf.ctypes.data, 1) == 0
    used = buf[buf[:, 5] > 0]
    lead = used[used[:, 2] > 0]      # CTAs whose MMA warp ran (2-SM: leaders only)
    tiles = max(used[:, 8].mean(), 1)
    m = used.mean(0)
    ml = lead.mean(0) if len(lead) else m
    print(f"{name}: {len(used)} CTAs, {tiles:.0f} tiles/CTA, kernel {m[5]:.0f} clk = {m[5]/tiles:.0f} clk/tile")
    print(f"   producer: total {m[0]/tiles:6.0f}/tile   blocked on empty {m[1]/tiles:6.0f}")
    print(f"   mma     : total {ml[2]/tiles:6.0f}/tile   blocked on tempty {ml[3]/tiles:6.0f}   on full {ml[4]/tiles:6.0f}")
    for ew in range(2):
        print(f"   epi w{ew}  : blocked on tfull {m[6+3*ew]/tiles:6.0f}/tile   busy after tfull {m[7+3*ew]/tiles:6.0f}")
    print(f"   epi w0 phases per tile: tcgen05.ld+wait {m[12]/tiles:6.0f}   math {m[13]/tiles:6.0f}   staging+stores {m[14]/tiles:6.0f}")


if __name__ == "__main__":
    assert fn(None, 1) == 0
    for n in sys.argv[1:] or ["pi", "qkv", "geglu", "po", "ffout"]:
        S.SHAPES[n]()
        dump(n)
