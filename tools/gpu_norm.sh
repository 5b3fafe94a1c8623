#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/kernel_check.py norm elementwise > gpurun_out/kernel_check.log 2>&1; grep -v PASS gpurun_out/kernel_check.log | tail
timeout 300 python tools/norm_bench.py 2>&1 | tee gpurun_out/norm_bench.log
