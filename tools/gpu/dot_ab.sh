#!/bin/bash
# A/B of DotGatherKernel's lane walk on the m=20 slice's last step (unit ttgt:DotGatherKernel in the profile)
cd "$(dirname "$0")/../.."
for v in 0 1 2 3 4 5; do
  echo "JB_DOT_LANE_A_BITS=$v"
  JB_DOT_LANE_A_BITS=$v python tools/plan_profile.py sycamore53_m20 0 --top 3 2>/dev/null | grep -E "total ms|DotGather"
done
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "final_dot" 2>&1 | tail -2
