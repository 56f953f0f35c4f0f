import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_micro, bench
peak,_=bench.measured_peaks()
res=[]
c=int(sys.argv[1]) if len(sys.argv)>1 else 4
r=26
rng = np.random.default_rng(100 * r + c)
ia = list(range(r)); common = sorted(rng.choice(r, c, replace=False).tolist())
ib = common + list(range(100, 100 + c)); ib = [ib[i] for i in rng.permutation(2 * c)]
print("common", common, "ib", ib)
bench_micro.run_contract(np.complex64, r, ia, 2 * c, ib, 3, peak, "S1_skinny", res, check=False)
