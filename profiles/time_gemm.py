import sys; sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import xerus_b200 as xb
from xerus_b200._lib import call
xb.init(0)
stream=torch.cuda.ExternalStream(xb.stream_handle())
for (m,n,k) in [(4096,4096,4096),(1024,8192,512),(8192,512,1024),(2048,2048,2048),(512,512,512),(256,256,256)]:
    A=torch.randn(m,k,dtype=torch.float64,device='cuda'); B=torch.randn(k,n,dtype=torch.float64,device='cuda'); Cm=torch.empty(m,n,dtype=torch.float64,device='cuda')
    torch.cuda.synchronize()
    def run(): call("xb_dev_gemm", Cm.data_ptr(), n, m, n, 1.0, A.data_ptr(), k, 0, k, B.data_ptr(), n, 0, 0.0)
    run(); xb.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    reps=5
    e0.record(stream)
    for _ in range(reps): run()
    e1.record(stream); xb.synchronize()
    ms=e0.elapsed_time(e1)/reps
    ref=A@B; torch.cuda.synchronize()
    s0,s1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(reps): ref=A@B
    s1.record(); torch.cuda.synchronize()
    ms2=s0.elapsed_time(s1)/reps
    err=((Cm-ref).norm()/ref.norm()).item()
    print(m,n,k,'xb %.3f ms %.1f TF | cublas %.3f ms %.1f TF | err %.1e'%(ms,2*m*n*k/ms/1e9,ms2,2*m*n*k/ms2/1e9,err),flush=True)
