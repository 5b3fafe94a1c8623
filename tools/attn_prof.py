"""One attention launch (self-attention, N=4096, 5 heads, batch 64) for ncu captures."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gemm_bench as g  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    nkv = int(sys.argv[2]) if len(sys.argv) > 2 else n
    g.bench_attn(64, 5, n, nkv)
