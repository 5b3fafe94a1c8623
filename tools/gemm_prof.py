"""One grouped-GEMM shape for ncu captures: python tools/gemm_prof.py M K N bn [res] [geglu]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gemm_bench as g  # noqa: E402

if __name__ == "__main__":
    M, K, N, bn = [int(x) for x in sys.argv[1:5]]
    res = len(sys.argv) > 5 and sys.argv[5] == "1"
    geglu = len(sys.argv) > 6 and sys.argv[6] == "1"
    g.bench_linear(M, K, N, bn, residual=res, geglu=geglu)
