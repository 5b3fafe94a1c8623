#!/bin/bash
# round 2, quick perf visit: selected kernel checks, U-Net parity, forward bench with per-launch profile, ncu launch list
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
timeout 600 python tools/kernel_check.py ${KCHECK:-gemm_ln_fold gemm_res groupnorm stream} > gpurun_out/kernel_check.log 2>&1; grep -v PASS gpurun_out/kernel_check.log | tail -10
timeout 1200 python -m pytest tests/test_unet_gpu.py -m gpu -q -x > gpurun_out/pytest_unet.log 2>&1; tail -4 gpurun_out/pytest_unet.log
APTP_PROFILE_DUMP=gpurun_out/kernel_profile.tsv timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-secondary > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
for k in ('value','ms_per_step','gemm_tflops','gemm_ms','attention_tflops','attention_ms','hbm_kernels_gbs','hbm_kernels_ms','step_frac_of_peak_kept_work','clocks'):
    print(k, d.get(k))
PY
tail -3 gpurun_out/bench.err
if [ -n "$NCU_LIST" ]; then
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --no-secondary > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv | head -40
fi
