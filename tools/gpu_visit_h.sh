#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_backward_gpu.py tests/test_sampling_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
APTP_PROFILE_DUMP=gpurun_out/kernel_profile.tsv timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'gemm', d['roofline']['achieved'], d['roofline']['kernel_ms_per_step'], 'attn', d['roofline']['attention'], d['clocks'])
PY
