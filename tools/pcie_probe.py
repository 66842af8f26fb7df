import torch, time
n=1920*1080
h=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device='cuda'); h2=torch.empty(n,dtype=torch.uint8).pin_memory()
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def t(fn,reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps
dt=t(lambda: d.copy_(h,non_blocking=True)); print(f"H2D 2MB: {dt*1e6:.1f} us  {n/dt/1e9:.1f} GB/s")
dt=t(lambda: h2.copy_(d,non_blocking=True)); print(f"D2H 2MB: {dt*1e6:.1f} us  {n/dt/1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d,non_blocking=True)
dt=t(both); print(f"H2D+D2H concurrent: {dt*1e6:.1f} us per pair")
big=torch.empty(64*n,dtype=torch.uint8).pin_memory(); dbig=torch.empty(64*n,dtype=torch.uint8,device='cuda')
dt=t(lambda: dbig.copy_(big,non_blocking=True),20); print(f"H2D 133MB: {64*n/dt/1e9:.1f} GB/s")
