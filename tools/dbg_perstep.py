import json, os, sys
import numpy as np
ROOT="/root/repo"
sys.path.insert(0, ROOT)
from jet_b200 import ContractionPlan, NetworkFile
from jet_b200.slicing import find_slices
from oracle import ref
DATA=os.path.join(ROOT,"data")
stem="sycamore53_m20"
meta=json.load(open(os.path.join(DATA,stem+".meta.json"))); js=json.load(open(os.path.join(DATA,stem+".json")))
leaf=[t[1] for t in js["tensors"]]; dims={i:2 for idx in leaf for i in idx}; path=[tuple(p) for p in js["path"]]
full=find_slices(leaf,dims,path,list(meta["sliced_indices"]),extra=12)
net=NetworkFile.load(os.path.join(DATA,stem+".json"),np.complex64)
with ContractionPlan(net, full, keep_intermediates=True) as plan:
    plan.reset(); plan.run(12345,1); plan.sync()
    n=0
    for st in plan.steps():
        ma, mb = plan.node_modes(st.node_a), plan.node_modes(st.node_b)
        if st.m*st.n*st.k > 2**27 or max(st.m*st.k, st.k*st.n, st.m*st.n) > 2**24 or not ma or not mb: continue
        a=plan.node(st.node_a).reshape(plan.node_shape(st.node_a)); b=plan.node(st.node_b).reshape(plan.node_shape(st.node_b))
        try:
            ref.contract(ma,a,mb,b)
        except Exception as e:
            n+=1
            if n<=3: print("FAIL", st.node_a, st.node_b, ma, a.shape, mb, b.shape, e)
        else:
            if n<3: print("ok", ma, a.shape, mb, b.shape)
    print("failures", n)
