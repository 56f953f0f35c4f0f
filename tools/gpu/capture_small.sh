#!/bin/bash
# ncu --set full captures of the two round-2 gather kernels, exported to CSV on the box:
#   DotGatherKernel<float,4,4>: the last step of an m=20 slice; SmallGemmGatherKernel<float>: the batched 16 x 16 x 4096
#   step of m10 s=10 (64 slices per launch, clusters of 8 CTAs).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:DotGather -c 1 -o /tmp/cap_dot -f python bench.py --no-graph --lanes 1 --slices-per-step 1 --steps 1 --warmup 0 --no-cpu --no-others --strong-slices 0 > gpurun_out/cap_dot.log 2>&1
ncu -i /tmp/cap_dot.ncu-rep --page raw --csv > gpurun_out/cap_dot_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:SmallGemmGather -s 2 -c 1 -o /tmp/cap_sg -f python bench.py --no-graph --workload sycamore53_m10_s10 --lanes 1 --slices-per-step 256 --steps 1 --warmup 1 --no-cpu --no-others --strong-slices 0 > gpurun_out/cap_sg.log 2>&1
ncu -i /tmp/cap_sg.ncu-rep --page raw --csv > gpurun_out/cap_sg_raw.csv 2>/dev/null
ls -la gpurun_out/ | tail -6; tail -2 gpurun_out/cap_sg.log
