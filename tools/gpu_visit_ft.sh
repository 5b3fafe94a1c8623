#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -x -k "stride2 or col_sum" > gpurun_out/pytest_s2.log 2>&1; tail -5 gpurun_out/pytest_s2.log
timeout 600 python tools/finetune_check.py > gpurun_out/finetune_check.log 2>&1; tail -48 gpurun_out/finetune_check.log
