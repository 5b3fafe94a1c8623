#!/bin/bash
# stock attention: parity, microbench, one ncu --set full capture of the N=4096 self-attention launch
mkdir -p gpurun_out
timeout 300 python tools/kernel_check.py attention 2>&1 | tail -12
timeout 300 python tools/attn_prof.py 4096 4096
timeout 300 python tools/attn_prof.py 1024 1024
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 5 -c 1 -f -o gpurun_out/attn_new python tools/attn_prof.py 4096 4096 > gpurun_out/attn_ncu.log 2>&1; echo "ncu rc=$?"
