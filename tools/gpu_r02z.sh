#!/bin/bash
# round 2, final records on one B200: GPU test suite, smoke, image-list harness, both bench arms (timed), ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/z_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 Src/goofy_bench --images oracle/_ref/test-data --csv gpurun_out/z_images.csv > gpurun_out/z_images.txt 2> gpurun_out/z_images.err; echo "harness images rc=$?"; tail -1 gpurun_out/z_images.txt | cut -c1-900
Src/goofy_bench --size 8192 --textures 4 --iters 20 --rgb24 > gpurun_out/z_harness_synth.json 2>&1
SECONDS=0
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/z_bench_ref_n1.json 2> gpurun_out/z_bench_ref_n1.err; echo "reference arm: $SECONDS s"
SECONDS=0
python bench.py > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench_n1.err; echo "bench.py (defaults): $SECONDS s"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/z_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/z_bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/z_bench_ref_n1.json','gpurun_out/z_bench_n1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['metric'], round(d['value']), 'MP/s', 'e2e', d.get('e2e',{}).get('value'))
        for k in ('roofline','clocks','per_texture_launch','other_codec','dual_output','rgb24_input','e2e','cpu_baseline'):
            if k in d: print('  ',k, json.dumps(d[k])[:900])
    except Exception as e: print(f,'ERR',e)
PY
