#!/bin/bash
# ncu evidence for one bench step (run after gpu_round.sh in the same visit):
#  1. DRAM bytes + tensor-pipe activity of EVERY launch of the step (cheap metrics, few passes)
#  2. one --set full capture of representative grouped-GEMM and attention launches
mkdir -p gpurun_out
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --clock-control none \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --csv --log-file gpurun_out/launch_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_metrics.log 2>&1; echo "ncu metrics rc=$?"
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:grouped_gemm_kernel -s ${GEMM_SKIP:-20} -c ${GEMM_COUNT:-12} -f -o gpurun_out/gemm_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm_full.log 2>&1; echo "ncu gemm full rc=$?"
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:attention_kernel -c 4 -f -o gpurun_out/attn_full \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn_full.log 2>&1; echo "ncu attn full rc=$?"
ls -la gpurun_out
