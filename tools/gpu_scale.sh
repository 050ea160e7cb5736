#!/bin/bash
# N-GPU scaling run under torchrun: bench (B200 arm + reference arm); add "full" for configs 3/4 and the C++ harness
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 5 --warmup 2 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N rc=$?"
if [ "$2" = "full" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/bench_configs.py --config batch1024 > gpurun_out/cfg_batch1024_n$N.json 2> gpurun_out/cfg_batch1024_n$N.err; echo "batch1024 N=$N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 tools/bench_configs.py --config strip16384 > gpurun_out/cfg_strip16384_n$N.json 2> gpurun_out/cfg_strip16384_n$N.err; echo "strip16384 N=$N rc=$?"
timeout 300 Src/goofy_bench --gpus $N --images $((N*2)) --iters 20 > gpurun_out/harness_n$N.json 2> gpurun_out/harness_n$N.err; echo "harness N=$N rc=$?"
fi
python - <<PY
import json
for f in ('bench_n$N','bench_ref_n$N','cfg_batch1024_n$N','cfg_strip16384_n$N','harness_n$N'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        keys={k:d[k] for k in ('value','n_gpus','ms_per_step','results','e2e','dxt1','etc1','dual','multi_gpu_equals_single_gpu') if k in d}
        if 'config' in d: keys['cpu_binding']=d['config'].get('cpu_binding')
        print(f, json.dumps(keys)[:700])
    except Exception as e: print(f,'ERR',e)
PY
