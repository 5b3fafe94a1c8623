#!/bin/bash
# ncu --set full capture of the transformer GEMMs of the first 64x64 block (pi, qkv, o, q, kv, o, geglu, ff-out, po)
mkdir -p gpurun_out
APTP_CUDA_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:grouped_gemm_kernel -s ${GEMM_SKIP:-6} -c ${GEMM_COUNT:-9} -f -o gpurun_out/gemm_full_r02 \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --no-secondary > gpurun_out/ncu_gemm_full.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/gemm_full_r02.ncu-rep --page raw --csv > gpurun_out/gemm_full_r02_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/gemm_full_r02_raw.csv')))
h=rows[0]
want=['Kernel Name','gpu__time_duration.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__m_xbar2l1tex_read_bytes.sum','sm__inst_issued.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','sm__warps_active.avg.pct_of_peak_sustained_active']
idx=[(w,h.index(w)) for w in want if w in h]
for r in rows[2:]:
    print(' | '.join(f"{w.split('.')[0][-28:]}={r[i][:22]}" for w,i in idx))
PY
