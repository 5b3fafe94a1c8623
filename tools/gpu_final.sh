#!/bin/bash
# Round-final single-GPU visit: parity tests, all bench workloads (+ reference arm), ncu launch lists / metrics / full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
APTP_PROFILE_DUMP=gpurun_out/kernel_profile.tsv timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err; cut -c1-300 gpurun_out/bench_plain.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c1-300 gpurun_out/bench_train.json
timeout 600 python bench.py --workload sample --steps 2 --warmup 1 > gpurun_out/bench_sample.json 2> gpurun_out/bench_sample.err; cut -c1-300 gpurun_out/bench_sample.json
timeout 600 python bench.py --workload finetune --steps 3 --warmup 3 --train-batch 32 > gpurun_out/bench_finetune.json 2> gpurun_out/bench_finetune.err; cut -c1-300 gpurun_out/bench_finetune.json
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
bash tools/gpu_ncu.sh > gpurun_out/gpu_ncu.log 2>&1; tail -3 gpurun_out/gpu_ncu.log
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/train_launches.csv python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1; echo "ncu train rc=$?"
ls -la gpurun_out | head -40
