"""Pinned-memory copy rates of this box (2 / 4 / 64 MB, each direction and both at once): the ceiling of bench.py's e2e line."""
import torch,time
torch.cuda.set_device(0)
for mb in (2,4,64):
    n=mb*1024*1024//4
    h=torch.empty(n).pin_memory(); d=torch.empty(n,device='cuda'); h2=torch.empty(n).pin_memory(); d2=torch.empty(n,device='cuda')
    s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
    def run(f,reps=50):
        torch.cuda.synchronize(); t=time.perf_counter()
        for _ in range(reps): f()
        torch.cuda.synchronize(); return (time.perf_counter()-t)/reps
    t=run(lambda: d.copy_(h,non_blocking=True)); print(mb,'MB H2D',mb/1024/t,'GB/s', t*1e6,'us')
    t=run(lambda: h2.copy_(d2,non_blocking=True)); print(mb,'MB D2H',mb/1024/t,'GB/s')
    def both():
        with torch.cuda.stream(s1): d.copy_(h,non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
    t=run(both); print(mb,'MB both dirs',2*mb/1024/t,'GB/s total')
