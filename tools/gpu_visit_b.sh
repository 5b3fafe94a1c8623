#!/bin/bash
# Visit B: new kernels (dkv rewrite, loss + macs kernels) -> parity tests, bwd-attention microbench, train bench
mkdir -p gpurun_out
timeout 300 python tools/attn_bwd_bench.py > gpurun_out/attn_bwd_bench.log 2>&1; cat gpurun_out/attn_bwd_bench.log | tail -8
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c1-400 gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
