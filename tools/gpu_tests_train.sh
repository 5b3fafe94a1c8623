#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c1-330 gpurun_out/bench_train.json; tail -3 gpurun_out/bench_train.err
