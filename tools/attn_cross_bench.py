"""Cross-attention (77 keys) and self-attention microbench: python tools/attn_cross_bench.py"""
import sys
sys.path.insert(0, "tools")
import gemm_bench as g

g.bench_attn(64, 5, 4096, 77)
g.bench_attn(64, 10, 1024, 77)
g.bench_attn(64, 20, 256, 77)
g.bench_attn(64, 5, 4096, 4096)
