#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:StreamPacked -s 1 -c 1 -o gpurun_out/r2l_stream16 -f python tools/prof_stream16.py 4 > gpurun_out/r2l_ncu.log 2>&1; tail -3 gpurun_out/r2l_ncu.log
