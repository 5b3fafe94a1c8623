#!/bin/bash
# attention variants (tools/build_variant.sh): parity checks + microbench per variant
mkdir -p gpurun_out
for lib in "" variants/libaptp_*.so; do
  echo "=== ${lib:-stock}" 
  APTP_LIB=${lib:+$PWD/$lib} timeout 300 python tools/kernel_check.py attention 2>&1 | tail -8
  APTP_LIB=${lib:+$PWD/$lib} timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0, 'tools')
import gemm_bench as g
g.bench_attn(64, 5, 4096, 4096)
g.bench_attn(64, 10, 1024, 1024)
g.bench_attn(64, 20, 256, 256)
g.bench_attn(64, 5, 4096, 77)
PY
done 2>&1 | tee gpurun_out/attn_sweep.log
