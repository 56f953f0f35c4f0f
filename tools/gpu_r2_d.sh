#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tools/plan_profile.py sycamore53_m20_t31 0 --top 12 > gpurun_out/r2d_m20_t31_profile.txt 2>&1; head -24 gpurun_out/r2d_m20_t31_profile.txt
timeout 1200 python bench.py --no-cpu --steps 3 --warmup 1 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -5 gpurun_out/r2d_bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','tflops','e2e','strong','peaks')})
    print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline'].get('fma'))
    for o in d.get('other_workloads',[]): print(json.dumps(o)[:400])
except Exception as e: print('parse failed', e)
PY
