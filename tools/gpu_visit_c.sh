#!/bin/bash
# Visit C: bwd attention microbench + backward / train parity
mkdir -p gpurun_out
timeout 300 python tools/attn_bwd_bench.py > gpurun_out/attn_bwd_bench.log 2>&1; cat gpurun_out/attn_bwd_bench.log | tail -8
timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
