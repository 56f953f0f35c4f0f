#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 jet_b200/cpp/test_dropin data/_ref/m10.json > gpurun_out/r2b_dropin.log 2>&1; echo "dropin rc=$?" >> gpurun_out/r2b_dropin.log; tail -5 gpurun_out/r2b_dropin.log
timeout 900 python tools/plan_profile.py sycamore53_m20 0 > gpurun_out/r2b_m20_profile.txt 2>&1; cat gpurun_out/r2b_m20_profile.txt | head -60
timeout 600 python bench.py --workload sycamore53_m20 --no-cpu --steps 3 --warmup 1 --slices-per-step 4 --lanes 1 > gpurun_out/r2b_bench_m20.json 2> gpurun_out/r2b_bench_m20.err; tail -c 1500 gpurun_out/r2b_bench_m20.json; tail -3 gpurun_out/r2b_bench_m20.err
