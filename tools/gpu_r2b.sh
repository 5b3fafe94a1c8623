#!/bin/bash
# round 2, visit B: full GPU suite + the new bench line (library bar, secondary workloads, flat keys)
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
APTP_PROFILE_DUMP=gpurun_out/kernel_profile.tsv timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
for k in ('value','ms_per_step','e2e','gemm_tflops','gemm_ms','attention_tflops','attention_ms','hbm_kernels_gbs','hbm_kernels_ms','hbm_frac','step_frac_of_peak_kept_work','library_baseline','vs_library_mixed','vs_library_dense','secondary','cpu_baseline','clocks'):
    print(k, d.get(k))
PY
tail -5 gpurun_out/bench.err
