#!/bin/bash
cd "$(dirname "$0")/.."
timeout 300 python bench.py --workload sycamore53_m10_s6 --no-cpu --no-others --steps 4 --warmup 2 --slices-per-step 256 --lanes 1 --strong-slices 0 2>&1 | tail -5 | cut -c1-600
timeout 300 python bench.py --workload gbs_fock4_total10_s2 --no-cpu --no-others --steps 4 --warmup 2 --slices-per-step 256 --lanes 2 --strong-slices 0 2>&1 | tail -5 | cut -c1-600
