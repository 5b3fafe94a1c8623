#!/bin/bash
# Same-box A/B of engine variants: AB="VAR=a VAR=b ..." ; each setting is benched REPS times, interleaved (box-to-box clock
# differences of +-2 % otherwise hide sub-millisecond effects). Prints ms/step of the graph-replayed forward.
mkdir -p gpurun_out
: ${AB:="APTP_LN_FOLD=1 APTP_LN_FOLD=0"}
: ${REPS:=2}
for r in $(seq $REPS); do
  for v in $AB; do
    env $v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-library-baseline --no-secondary > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
    python - "$v" "$r" <<'PY'
import json,sys
d=json.load(open('gpurun_out/ab.json'))
print(f"AB {sys.argv[1]:28s} rep{sys.argv[2]} ms/step={d['ms_per_step']:.3f} gemm_ms={d.get('gemm_ms')} attn_ms={d.get('attention_ms')} hbm_ms={d.get('hbm_kernels_ms')} clk={d['clocks'].get('sm_mhz')}", flush=True)
PY
  done
done 2>&1 | tee gpurun_out/ab.log
