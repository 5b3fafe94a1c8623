#!/bin/bash
# ncu --set full of one shortk_bench shape per library variant: VARIANTS="stock late" SHAPE=pi bash tools/gpu_ncu_shortk.sh
mkdir -p gpurun_out
for v in ${VARIANTS:-stock}; do
  lib=""; [ "$v" != "stock" ] && lib="variants/libaptp_$v.so"
  APTP_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:grouped_gemm_kernel -s 5 -c 1 -f \
    -o gpurun_out/shortk_${SHAPE:-pi}_$v python tools/shortk_bench.py ${SHAPE:-pi} > gpurun_out/ncu_shortk_$v.log 2>&1; echo "ncu $v rc=$?"
  ncu -i gpurun_out/shortk_${SHAPE:-pi}_$v.ncu-rep --page raw --csv > gpurun_out/shortk_${SHAPE:-pi}_${v}_raw.csv 2>/dev/null
done
