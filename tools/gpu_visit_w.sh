#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_backward_gpu.py -m gpu -q -x -k wgrad > gpurun_out/pytest_wgrad.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_wgrad.log
tail -25 gpurun_out/pytest_wgrad.log
timeout 300 python tools/wgrad_bench.py > gpurun_out/wgrad_bench.log 2>&1; tail -8 gpurun_out/wgrad_bench.log
