#!/bin/bash
# round-2 record run, part 2: racecheck on the barrier hand-off, full tests, default bench + reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_chain_gpu.py -m gpu -x -q > gpurun_out/r2q_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2q_racecheck.log; tail -3 gpurun_out/r2q_racecheck.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -4 gpurun_out/r2q_pytest.log
timeout 900 python tools/plan_profile.py sycamore53_m20 0 --top 3 2>&1 | sed -n 2,6p
free -g | head -2
timeout 1500 python bench.py > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err; tail -2 gpurun_out/r2q_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','tflops','e2e','clocks')}); print(d.get('cpu_baseline')); print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['fma']['frac'], d['strong'])
for o in d.get('other_workloads',[]): print(json.dumps(o)[:260])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err; tail -c 700 gpurun_out/r2q_bench_ref.json; tail -2 gpurun_out/r2q_bench_ref.err
