#!/bin/bash
# bench.py at N GPUs (headline + the named configs[3]/[4] + shard scheduler), and the reference arm
mkdir -p gpurun_out
N=${1:-1}
run() { if [ "$N" = "1" ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; fi; }
run bench.py --impl reference --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err || tail -5 gpurun_out/bench_ref_n$N.err
run bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -20 gpurun_out/bench_n$N.err
python - <<PY
import json
for f in ('gpurun_out/bench_ref_n$N.json','gpurun_out/bench_n$N.json'):
    try:
        d=json.load(open(f))
        print(f, d['metric'], round(d['value']), 'MP/s', 'e2e', d.get('e2e',{}).get('value'))
        for k in ('roofline','clocks','per_texture_launch','other_codec','dual_output','e2e','configs','sharded_api','cpu_baseline'):
            if k in d: print('  ',k, json.dumps(d[k])[:1500])
    except Exception as e: print(f,'ERR',e)
PY
