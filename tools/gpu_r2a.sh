#!/bin/bash
# round 2, visit A: new kernels (fp32 stream, deterministic GroupNorm stats) + whole-U-Net parity with logged margins + bench
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python tools/kernel_check.py > gpurun_out/kernel_check.log 2>&1; grep -v PASS gpurun_out/kernel_check.log | tail -30
timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -q > gpurun_out/pytest_unet.log 2>&1; tail -15 gpurun_out/pytest_unet.log
cat gpurun_out/parity_metrics.jsonl
APTP_PROFILE_DUMP=gpurun_out/kernel_profile.tsv timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
