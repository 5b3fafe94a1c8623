#!/bin/bash
# round 2, N-GPU visit (gpurun --gpus N): NCCL tests + the driver's torchrun launch of bench.py (forward line with the
# secondary train / sampling numbers) + the reference arm
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x > gpurun_out/pytest_multi_n$N.log 2>&1; tail -12 gpurun_out/pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/bench_n$N.err
python - "$N" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/bench_n{sys.argv[1]}.json').read().strip().splitlines()[-1])
for k in ('n_gpus','value','ms_per_step','e2e','secondary','clocks'):
    print(k, d.get(k))
PY
