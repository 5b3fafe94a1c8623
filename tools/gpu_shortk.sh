#!/bin/bash
# same-box comparison on the level-0 transformer GEMM shapes of kernel-tuning builds (VARIANTS="a b" -> variants/libaptp_a.so)
# and/or environment settings (ENVS="APTP_GEMM_1SM=1 ..."):  SHAPES="pi qkv" bash tools/gpu_shortk.sh
mkdir -p gpurun_out
: > gpurun_out/shortk.log
for rep in 1 2; do
  echo "== stock rep$rep" >> gpurun_out/shortk.log
  timeout 300 python tools/shortk_bench.py $SHAPES >> gpurun_out/shortk.log 2>&1
  for e in $ENVS; do
    echo "== $e rep$rep" >> gpurun_out/shortk.log
    env $e timeout 300 python tools/shortk_bench.py $SHAPES >> gpurun_out/shortk.log 2>&1
  done
  for v in $VARIANTS; do
    echo "== $v rep$rep" >> gpurun_out/shortk.log
    APTP_LIB=variants/libaptp_$v.so timeout 300 python tools/shortk_bench.py $SHAPES >> gpurun_out/shortk.log 2>&1
  done
done
cat gpurun_out/shortk.log
