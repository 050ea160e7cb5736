#!/bin/bash
# round 2, session S (N GPUs): bench line at N ranks: e2e with alpha-stripped staging (AUTO) and without, same run
N=${1:-2}
mkdir -p gpurun_out
nproc > gpurun_out/s_nproc_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-configs \
  > gpurun_out/s_bench_n$N.json 2> gpurun_out/s_bench_n$N.err
tail -2 gpurun_out/s_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/s_bench_n$N.json').read().strip().splitlines()[-1])
print('value', d['value'])
e=d['e2e']; print({k:v for k,v in e.items() if k not in ('pcie_bound_note','api')})
PY
