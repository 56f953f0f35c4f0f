import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from jet_b200 import ops
dev=torch.device('cuda:0')
for m,n,k in [(1<<21,64,64),(1<<18,64,64),(4096,4096,4096),(1<<16,256,256),(8192,64,1024)]:
    a=torch.randn(m*k,dtype=torch.complex128,device=dev); b=torch.randn(k*n,dtype=torch.complex128,device=dev); c=torch.empty(m*n,dtype=torch.complex128,device=dev)
    wsb=ops.gemm_ws_bytes(np.complex128,m,n,k); ws=torch.empty(max(wsb,16),dtype=torch.uint8,device=dev)
    f=lambda: ops.gemm_device(np.complex128,m,n,k,a.data_ptr(),b.data_ptr(),c.data_ptr(),ws.data_ptr(),wsb)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(m,n,k,"ms %.3f TFLOP/s %.1f"%(ms, 8*m*n*k/ms/1e9))
