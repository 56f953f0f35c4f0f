#!/bin/bash
# A/B of the ChainKernel group layouts on one B200: the isolated launch shape (tools/prof_chain.py) and the m=20 bench.
# (The DESIGN §3 K3 numbers for "4x128" come from this script.)
cd "$(dirname "$0")/../.."
m20() { python bench.py --no-others --no-cpu --strong-slices 0 --steps 3 --warmup 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['fma']['frac'])"; }
echo "default: two groups of 8 warps, 2^13-element tiles"; python tools/prof_chain.py 27 6 | tail -1; m20
echo "JB_CHAIN_LAYOUT=4x128: four groups of 4 warps, 2^12-element tiles"; JB_CHAIN_LAYOUT=4x128 python tools/prof_chain.py 27 6 | tail -1; JB_CHAIN_LAYOUT=4x128 m20
