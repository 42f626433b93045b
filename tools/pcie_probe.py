import torch, time
dev=torch.device("cuda",0)
for mb in (65.5, 28.8):
    n=int(mb*1e6/8)
    h=torch.empty(n,dtype=torch.float64).pin_memory(); d=torch.empty(n,dtype=torch.float64,device=dev)
    for direction in ("h2d","d2h"):
        torch.cuda.synchronize()
        t0=time.perf_counter()
        for _ in range(10):
            if direction=="h2d": d.copy_(h,non_blocking=True)
            else: h.copy_(d,non_blocking=True)
        torch.cuda.synchronize()
        dt=(time.perf_counter()-t0)/10
        print(direction, mb, "MB", round(dt*1e3,3), "ms", round(mb/1e3/dt,1), "GB/s")
# both directions at once
n1=int(65.5e6/8); n2=int(28.8e6/8)
h1=torch.empty(n1,dtype=torch.float64).pin_memory(); d1=torch.empty(n1,dtype=torch.float64,device=dev)
h2=torch.empty(n2,dtype=torch.float64).pin_memory(); d2=torch.empty(n2,dtype=torch.float64,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); print("both", round((time.perf_counter()-t0)/10*1e3,3),"ms")
