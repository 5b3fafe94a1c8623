#!/bin/bash
# attention kernel A/B on one box: parity checks + microbench, stock library vs variants/libaptp_*.so
mkdir -p gpurun_out
for lib in "" variants/libaptp_oldattn.so; do
  echo "=== ${lib:-new}"
  APTP_LIB=${lib:+$PWD/$lib} timeout 300 python tools/kernel_check.py attention 2>&1 | grep -v "^PASS" | tail -8
  APTP_LIB=${lib:+$PWD/$lib} timeout 300 python - <<'PY'
import sys
sys.path.insert(0, 'tools')
import gemm_bench as g
for _ in range(2):
    g.bench_attn(64, 5, 4096, 4096)
    g.bench_attn(64, 10, 1024, 1024)
    g.bench_attn(64, 20, 256, 256)
    g.bench_attn(64, 20, 64, 64)
    g.bench_attn(64, 5, 4096, 77)
    g.bench_attn(64, 10, 1024, 77)
    g.bench_attn(64, 20, 256, 77)
PY
done 2>&1 | tee gpurun_out/attn_ab.log
