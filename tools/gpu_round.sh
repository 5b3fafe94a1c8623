#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), ncu launch list of one step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
APTP_PROFILE_DUMP=gpurun_out/kernel_profile.tsv timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
