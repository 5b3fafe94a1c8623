#!/bin/bash
# round 2, final single-GPU evidence visit: GPU suite (x RUNS, stability), smoke(), full bench line, reference arm, the other
# workloads, ncu launch list + per-launch DRAM / tensor metrics, one --set full capture of GEMM (2-SM conv + K=320 linear) and
# attention launches. Outputs under gpurun_out/ (copied to profiles/r02_* afterwards).
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
for i in $(seq ${RUNS:-3}); do
  timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_run$i.log 2>&1; echo "pytest run$i rc=$?"; tail -1 gpurun_out/pytest_gpu_run$i.log
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
APTP_PROFILE_DUMP=gpurun_out/kernel_profile.tsv timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py --workload train --steps 5 --warmup 3 --no-secondary > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c1-200 gpurun_out/bench_train.json
timeout 600 python bench.py --workload sample --steps 2 --warmup 1 --no-secondary > gpurun_out/bench_sample.json 2> gpurun_out/bench_sample.err; cut -c1-200 gpurun_out/bench_sample.json
timeout 600 python bench.py --workload finetune --steps 3 --warmup 3 --train-batch 32 --no-secondary > gpurun_out/bench_finetune.json 2> gpurun_out/bench_finetune.err; cut -c1-200 gpurun_out/bench_finetune.json
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --clock-control none \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --csv --log-file gpurun_out/launch_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --no-secondary > gpurun_out/ncu_metrics.log 2>&1; echo "ncu metrics rc=$?"
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:grouped_gemm_kernel -s 4 -c 14 -f -o gpurun_out/gemm_full_r02 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --no-secondary > gpurun_out/ncu_gemm_full.log 2>&1; echo "ncu gemm full rc=$?"
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:attention_kernel -c 2 -f -o gpurun_out/attn_full_r02 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --no-secondary > gpurun_out/ncu_attn_full.log 2>&1; echo "ncu attn full rc=$?"
ncu -i gpurun_out/gemm_full_r02.ncu-rep --page raw --csv > gpurun_out/gemm_full_r02_raw.csv 2>/dev/null
ncu -i gpurun_out/attn_full_r02.ncu-rep --page raw --csv > gpurun_out/attn_full_r02_raw.csv 2>/dev/null
ls -la gpurun_out | wc -l
