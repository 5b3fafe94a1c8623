#!/bin/bash
# config 2 quick visit: backward / train parity tests + the train bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_unet_gpu.py -m gpu -q > gpurun_out/pytest_train.log 2>&1; tail -3 gpurun_out/pytest_train.log
timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c1-330 gpurun_out/bench_train.json; tail -2 gpurun_out/bench_train.err
