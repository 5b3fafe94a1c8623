#!/bin/bash
mkdir -p gpurun_out
APTP_CUDA_PROFILE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/finetune_launches.csv python bench.py --workload finetune --steps 1 --warmup 3 --train-batch 32 > gpurun_out/ncu_finetune.log 2>&1; echo "ncu finetune rc=$?"
