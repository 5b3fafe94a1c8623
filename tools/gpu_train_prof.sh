#!/bin/bash
# config 2 (pruning train step): bench line + ncu launch list of one step
mkdir -p gpurun_out
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cat gpurun_out/bench_train.json | cut -c1-600
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/train_launches.csv python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/ncu_train.log 2>&1; echo "ncu rc=$?"
