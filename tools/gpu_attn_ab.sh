#!/bin/bash
# attention kernel A/B on one box: parity checks + microbench, stock library vs VARIANTS="a b" (variants/libaptp_a.so)
mkdir -p gpurun_out
for v in stock $VARIANTS; do
  lib=""; [ "$v" != "stock" ] && lib="$PWD/variants/libaptp_$v.so"
  echo "=== $v"
  APTP_LIB=$lib timeout 90 python tools/kernel_check.py attention 2>&1 | grep -v "^PASS" | tail -4
  for r in 1 2; do APTP_LIB=$lib timeout 90 python tools/attn_cross_bench.py; done
done 2>&1 | tee gpurun_out/attn_ab.log
