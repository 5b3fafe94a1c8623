#!/bin/bash
# multi-GPU visit: bench lines at N GPUs (forward, train, routed sampling) through torchrun
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --workload sample --steps 2 --warmup 1 > gpurun_out/bench_sample_n$N.json 2> gpurun_out/bench_sample_n$N.err; cut -c1-900 gpurun_out/bench_sample_n$N.json; tail -3 gpurun_out/bench_sample_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_forward_n$N.json 2> gpurun_out/bench_forward_n$N.err; cut -c1-300 gpurun_out/bench_forward_n$N.json; tail -3 gpurun_out/bench_forward_n$N.err
timeout 600 $TR bench.py --gpus $N --workload train --steps 5 --warmup 3 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err; cut -c1-300 gpurun_out/bench_train_n$N.json; tail -3 gpurun_out/bench_train_n$N.err
