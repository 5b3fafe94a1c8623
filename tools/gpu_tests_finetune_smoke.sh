#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload finetune --steps 3 --warmup 3 --train-batch 32 > gpurun_out/bench_finetune.json 2> gpurun_out/bench_finetune.err; cut -c1-260 gpurun_out/bench_finetune.json; tail -4 gpurun_out/bench_finetune.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
