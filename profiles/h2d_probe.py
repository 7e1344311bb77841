import torch, time
n=40_000_000
h=torch.empty(n,dtype=torch.float64).pin_memory()
d=torch.empty(n,dtype=torch.float64,device='cuda')
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t_copy(reps=5):
    torch.cuda.synchronize(); t=time.perf_counter()
    with torch.cuda.stream(s1):
        for _ in range(reps): d.copy_(h,non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter()-t)/reps
print("H2D alone GB/s", n*8/t_copy()/1e9)
a=torch.randn(200_000_000,device='cuda',dtype=torch.float64); b=torch.empty_like(a)
def busy():
    with torch.cuda.stream(s2):
        for _ in range(40): torch.sin(a,out=b)
torch.cuda.synchronize(); t=time.perf_counter(); busy(); torch.cuda.synchronize(); print("busy alone ms",(time.perf_counter()-t)*1e3)
torch.cuda.synchronize(); t=time.perf_counter(); busy()
with torch.cuda.stream(s1):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(s1)
    for _ in range(5): d.copy_(h,non_blocking=True)
    e1.record(s1)
torch.cuda.synchronize(); print("both ms",(time.perf_counter()-t)*1e3, "H2D during busy GB/s", 5*n*8/(e0.elapsed_time(e1)*1e-3)/1e9)
