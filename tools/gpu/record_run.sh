#!/bin/bash
# The round-2 record run on one B200 (through gpurun): tests, compute-sanitizer, ncu launch list + captures, default
# bench and the reference arm.  Outputs land in gpurun_out/ and are copied into profiles/ by hand (profiles/README.md).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/rec_pytest.log 2>&1; tail -4 gpurun_out/rec_pytest.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_chain_gpu.py tests/test_plan_gpu.py -m gpu -x -q -k "not m12 and not large" > gpurun_out/rec_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/rec_memcheck.log; tail -3 gpurun_out/rec_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_chain_gpu.py -m gpu -x -q > gpurun_out/rec_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/rec_racecheck.log; tail -3 gpurun_out/rec_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "final_dot or dot_and_gemv or dmma or small_output" > gpurun_out/rec_memcheck_gemm.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/rec_memcheck_gemm.log; tail -3 gpurun_out/rec_memcheck_gemm.log
# launch list of one bench step (graph off: ncu cannot replay the graph's kernel nodes; probes and other workloads off)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/rec_launches_m20.csv python bench.py --no-graph --lanes 1 --slices-per-step 1 --steps 1 --warmup 1 --no-cpu --no-others --strong-slices 0 > gpurun_out/rec_ncu_bench.log 2>&1
# full captures: the heaviest per-slice chain launches and the final dot
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ChainKernel -s 176 -c 30 -o gpurun_out/rec_chain_m20 -f python bench.py --no-graph --lanes 1 --slices-per-step 1 --steps 1 --warmup 0 --no-cpu --no-others --strong-slices 0 > gpurun_out/rec_ncu_chain.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:DotGather -c 1 -o gpurun_out/rec_dot_m20 -f python bench.py --no-graph --lanes 1 --slices-per-step 1 --steps 1 --warmup 0 --no-cpu --no-others --strong-slices 0 > gpurun_out/rec_ncu_dot.log 2>&1
timeout 1500 python bench.py > gpurun_out/rec_bench_n1.json 2> gpurun_out/rec_bench_n1.err; tail -2 gpurun_out/rec_bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/rec_bench_ref.json 2> gpurun_out/rec_bench_ref.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/rec_bench_n1.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'tflops', 'e2e', 'clocks')}); print(d.get('cpu_baseline'))
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['fma']['frac'], d['strong'])
for o in d.get('other_workloads', []):
    print(json.dumps(o)[:260])
print(open('gpurun_out/rec_bench_ref.json').read()[-500:])
PY
