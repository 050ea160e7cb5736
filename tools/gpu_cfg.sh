#!/bin/bash
mkdir -p gpurun_out
N=${1:-1}
run() { if [ "$N" = "1" ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; fi; }
run tools/bench_configs.py --config batch1024 > gpurun_out/cfg_batch1024_n$N.json 2> gpurun_out/cfg_batch1024_n$N.err || tail -5 gpurun_out/cfg_batch1024_n$N.err
run tools/bench_configs.py --config strip16384 > gpurun_out/cfg_strip16384_n$N.json 2> gpurun_out/cfg_strip16384_n$N.err || tail -5 gpurun_out/cfg_strip16384_n$N.err
run tools/bench_configs.py --config strip16384 --load-path tma > gpurun_out/cfg_strip16384_tma_n$N.json 2> gpurun_out/cfg_strip16384_tma_n$N.err || tail -5 gpurun_out/cfg_strip16384_tma_n$N.err
run tools/bench_configs.py --config batch1024 --load-path tma > gpurun_out/cfg_batch1024_tma_n$N.json 2> gpurun_out/cfg_batch1024_tma_n$N.err || tail -5 gpurun_out/cfg_batch1024_tma_n$N.err
if [ "$N" != "1" ]; then run bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err; fi
for f in gpurun_out/cfg_*_n$N.json gpurun_out/bench_n$N.json; do [ -f $f ] && python - <<PY
import json
try:
    d=json.load(open('$f'))
    print('$f', round(d['value']), 'MP/s', d.get('results') or {k:d[k] for k in ('roofline','e2e') if k in d})
except Exception as e: print('$f', 'ERR', e)
PY
done
