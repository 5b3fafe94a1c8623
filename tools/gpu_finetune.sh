#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -m gpu -q -x > gpurun_out/pytest_train.log 2>&1; tail -12 gpurun_out/pytest_train.log
timeout 900 python bench.py --workload finetune --steps 3 --warmup 3 --train-batch ${FT_BATCH:-16} > gpurun_out/bench_finetune.json 2> gpurun_out/bench_finetune.err; cut -c1-700 gpurun_out/bench_finetune.json; tail -6 gpurun_out/bench_finetune.err
