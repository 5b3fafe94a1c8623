#!/bin/bash
# N-GPU visit (gpurun --gpus N): NCCL tests + the driver's torchrun launch of bench.py (forward, train, reference arm)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_sampling_gpu.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
for w in forward train; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 --workload $w > gpurun_out/bench_${w}_n$N.json 2> gpurun_out/bench_${w}_n$N.err
  echo "rc=$?"; tail -1 gpurun_out/bench_${w}_n$N.json | cut -c1-400
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; tail -1 gpurun_out/bench_ref_n$N.json | cut -c1-300
