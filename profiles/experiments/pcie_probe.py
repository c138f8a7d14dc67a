import torch, time
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
for mb in (1, 4, 7.6, 16, 64, 256):
    n=int(mb*1e6/4)
    d=torch.empty(n,dtype=torch.float32,device="cuda"); h=torch.empty(n,dtype=torch.float32).pin_memory()
    for name,fn in (("D2H",lambda: h.copy_(d,non_blocking=True)),("H2D",lambda: d.copy_(h,non_blocking=True))):
        fn(); torch.cuda.synchronize(); e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/10
        print(f"{name} {mb} MB: {ms:.3f} ms  {mb*1e6/ms/1e6:.1f} GB/s")
# two D2H copies concurrently on two streams
n=int(3.8e6/4)
d1=torch.empty(n,device="cuda"); d2=torch.empty(n,device="cuda"); h1=torch.empty(n).pin_memory(); h2=torch.empty(n).pin_memory()
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): h1.copy_(d1,non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2,non_blocking=True)
torch.cuda.synchronize(); print("2 streams x 3.8 MB D2H:", (time.perf_counter()-t0)/20*1e3, "ms per pair")
import subprocess
print(subprocess.run("nvidia-smi -q | grep -i -A3 'PCIe Generation\\|Link Width' | head -20; nvidia-smi topo -m | head -8; nproc; numactl -H 2>/dev/null | head -5", shell=True, capture_output=True, text=True).stdout)
