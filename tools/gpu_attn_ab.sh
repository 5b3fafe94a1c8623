#!/bin/bash
# attention kernel A/B on one box: parity checks + microbench, stock library vs VARIANTS="a b" (variants/libaptp_a.so)
mkdir -p gpurun_out
for v in stock $VARIANTS; do
  lib=""; [ "$v" != "stock" ] && lib="$PWD/variants/libaptp_$v.so"
  echo "=== $v"
  APTP_LIB=$lib timeout 90 python tools/kernel_check.py attention 2>&1 | grep -v "^PASS" | tail -4
  APTP_LIB=$lib timeout 90 python - <<'PY'
import sys
sys.path.insert(0, 'tools')
import gemm_bench as g
for _ in range(2):
    g.bench_attn(64, 5, 4096, 4096)
    g.bench_attn(64, 10, 1024, 1024)
    g.bench_attn(64, 20, 256, 256)
    g.bench_attn(64, 5, 4096, 77)
    g.bench_attn(64, 10, 1024, 77)
PY
done 2>&1 | tee gpurun_out/attn_ab.log
