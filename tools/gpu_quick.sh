#!/bin/bash
# quick GPU visit: kernel checks + microbench (+ optional bench)
mkdir -p gpurun_out
timeout 600 python tools/kernel_check.py "$@" > gpurun_out/kernel_check.log 2>&1; grep -v PASS gpurun_out/kernel_check.log | tail -30
timeout 600 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
