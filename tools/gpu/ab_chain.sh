m20() { python bench.py --no-others --no-cpu --strong-slices 0 --steps 3 --warmup 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['fma']['frac'])"; }
echo "A: 2x256 const-steps on"; python tools/prof_chain.py 27 6 | tail -1; m20
echo "A: 2x256 const-steps off"; JB_CHAIN_CONST_STEPS=0 python tools/prof_chain.py 27 6 | tail -1; JB_CHAIN_CONST_STEPS=0 m20
echo "B: 4x128 const-steps off"; JB_CHAIN_LAYOUT=4x128 JB_CHAIN_CONST_STEPS=0 python tools/prof_chain.py 27 6 | tail -1; JB_CHAIN_LAYOUT=4x128 JB_CHAIN_CONST_STEPS=0 m20
python -m pytest tests/test_chain_gpu.py tests/test_plan_gpu.py -m gpu -x -q 2>&1 | tail -2
