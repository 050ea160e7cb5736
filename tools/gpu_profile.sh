#!/bin/bash
# ncu session: launch list of the bench command + one full capture of each encoder kernel + plain bench runs.
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --codec etc1 --no-cpu-baseline > gpurun_out/bench_etc1.json 2>> gpurun_out/bench.err; echo "bench etc1 rc=$?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "launchlist rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:encode_ -s 9 -c 3 -f -o gpurun_out/prof \
    python tools/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json
for f in ('bench','bench_etc1','bench_ref'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, round(d['value']), d.get('roofline',{}).get('achieved'), d.get('roofline',{}).get('frac'), 'e2e', d.get('e2e',{}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
