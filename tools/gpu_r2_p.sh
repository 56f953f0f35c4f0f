#!/bin/bash
# round-2 record run (1 GPU): tests, sanitizer, ncu launch list + full capture of the dominant launch, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; tail -5 gpurun_out/r2p_pytest.log
# compute-sanitizer on the final chain.cu / engine.cu / contract.cu (memcheck over chain + plan + batching tests, racecheck over chains)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_chain_gpu.py tests/test_plan_gpu.py -m gpu -x -q -k "not m12 and not large" > gpurun_out/r2p_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2p_memcheck.log; tail -4 gpurun_out/r2p_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_chain_gpu.py -m gpu -x -q > gpurun_out/r2p_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2p_racecheck.log; tail -4 gpurun_out/r2p_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "final_dot or dot_and_gemv or dmma" > gpurun_out/r2p_memcheck_gemm.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2p_memcheck_gemm.log; tail -3 gpurun_out/r2p_memcheck_gemm.log
# launch list of the bench command (graph off: ncu cannot replay the graph's kernel nodes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches_m20.csv python bench.py --no-graph --lanes 1 --slices-per-step 1 --steps 1 --warmup 1 --no-cpu --no-others --strong-slices 0 > gpurun_out/r2p_ncu_bench.log 2>&1; tail -2 gpurun_out/r2p_ncu_bench.log | cut -c1-200
# full capture of the largest chain launches of one m=20 slice
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ChainKernel -s 60 -c 3 -o gpurun_out/r2p_chain_m20 -f python bench.py --no-graph --lanes 1 --slices-per-step 1 --steps 1 --warmup 1 --no-cpu --no-others --strong-slices 0 > gpurun_out/r2p_ncu_chain.log 2>&1; tail -2 gpurun_out/r2p_ncu_chain.log | cut -c1-200
timeout 900 ncu --set full --clock-control none -k regex:DotGather -c 1 -o gpurun_out/r2p_dot_m20 -f python bench.py --no-graph --lanes 1 --slices-per-step 1 --steps 1 --warmup 0 --no-cpu --no-others --strong-slices 0 > gpurun_out/r2p_ncu_dot.log 2>&1
# the default bench, as the driver runs it
timeout 1500 python bench.py > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; tail -2 gpurun_out/r2p_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','tflops','e2e','clocks')}); print(d.get('cpu_baseline')); print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['fma']['frac'], d['strong'])
for o in d.get('other_workloads',[]): print(json.dumps(o)[:260])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2p_bench_ref.json 2> gpurun_out/r2p_bench_ref.err; tail -c 600 gpurun_out/r2p_bench_ref.json
