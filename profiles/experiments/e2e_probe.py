import sys, time, json
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import bench
from mars_ode_physics_b200 import Arena, capi
W=65536
ar=Arena(W); stream=torch.cuda.current_stream(); ar.set_stream(stream.cuda_stream)
sc=bench.build_rover_worlds(ar,W); ar.set_order_mode(1)
for _ in range(600): ar.collide_step(0.001)
ar.sync()
nj=4
cmds=[torch.empty((nj,W),dtype=torch.float32).pin_memory() for _ in range(4)]
for i,c in enumerate(cmds): c[:]=torch.from_numpy(sc["vel"])[None,:]+0.001*i
obs=[torch.empty((13,W),dtype=torch.float32).pin_memory() for _ in range(2)]
jobs=[torch.empty((nj,4,W),dtype=torch.float32).pin_memory() for _ in range(2)]
param=capi.PARAM_VEL|capi.PARAM_GROUP2
def run(K, mode):
    pending=None
    torch.cuda.synchronize(); t0=time.perf_counter()
    for k in range(K):
        if mode!="noh2d": ar.joints_set_param_batch_ptr(param, cmds[k%4].data_ptr(), nj)
        ar.collide_step(0.001)
        if mode=="full":
            t=ar.observe_begin(sc["chassis"], obs[k&1].data_ptr(), jobs[k&1].data_ptr())
        elif mode in ("body","noh2d"):
            t=ar.observe_begin(sc["chassis"], obs[k&1].data_ptr(), None)
        else: t=None
        if pending is not None: ar.io_wait(pending)
        pending=t
    if pending is not None: ar.io_wait(pending)
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/K*1e3
for mode in ("full","body","none","noh2d","full"):
    run(20,mode); print(mode, round(run(200,mode),4), "ms/step", flush=True)
# raw D2H timing
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
big=torch.empty((29,W),dtype=torch.float32,device="cuda"); hb=torch.empty((29,W),dtype=torch.float32).pin_memory()
torch.cuda.synchronize(); e0.record(); 
for _ in range(20): hb.copy_(big,non_blocking=True)
e1.record(); torch.cuda.synchronize(); print("D2H 7.6MB ms", e0.elapsed_time(e1)/20)
