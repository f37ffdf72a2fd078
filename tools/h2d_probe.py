import torch, time
dev=torch.device('cuda:0')
for mb in (26, 94, 256):
    n=mb*1024*1024//4
    h=torch.empty(n, dtype=torch.float32).pin_memory(); d=torch.empty(n, dtype=torch.float32, device=dev)
    s=torch.cuda.Stream()
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/10
    print("H2D %d MB: %.3f ms  %.1f GB/s"%(mb, ms, n*4/ms/1e6))
# two streams concurrently
n=47*1024*1024//4
h1=torch.empty(n).pin_memory(); h2=torch.empty(n).pin_memory(); d1=torch.empty(n,device=dev); d2=torch.empty(n,device=dev)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/10
print("2 streams x 47 MB: %.3f ms %.1f GB/s"%(dt*1e3, 2*n*4/dt/1e9))
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv", shell=True, capture_output=True, text=True).stdout)
