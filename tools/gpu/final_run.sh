#!/bin/bash
# Last check of the round on one B200: the whole GPU suite, smoke(), the default bench line and the reference arm.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/fin_pytest.log 2>&1; tail -3 gpurun_out/fin_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/fin_bench_n1.json 2> gpurun_out/fin_bench_n1.err; tail -2 gpurun_out/fin_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/fin_bench_ref.json 2> gpurun_out/fin_bench_ref.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/fin_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'tflops', 'e2e', 'clocks')}); print(d.get('cpu_baseline', {}).get('value'))
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['fma']['frac'], json.dumps(d['roofline']['other_kernels'])[:400])
for o in d.get('other_workloads', []):
    print(json.dumps(o)[:200])
print(open('gpurun_out/fin_bench_ref.json').read()[:300])
PY
