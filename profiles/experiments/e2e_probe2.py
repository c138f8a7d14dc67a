import sys, time, os
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import bench
from mars_ode_physics_b200 import Arena, capi
W=65536
def make(graph=True):
    ar=Arena(W); stream=torch.cuda.current_stream(); ar.set_stream(stream.cuda_stream)
    sc=bench.build_rover_worlds(ar,W); ar.set_order_mode(1)
    if not graph: ar.set_graph_launch(False)
    for _ in range(600): ar.collide_step(0.001)
    ar.sync()
    return ar, sc
nj=4
param=capi.PARAM_VEL|capi.PARAM_GROUP2
def run(ar, sc, K, mode, cmds, obs, jobs):
    pending=None
    torch.cuda.synchronize(); t0=time.perf_counter()
    for k in range(K):
        if mode in ("full","h2d_io"): ar.joints_set_param_batch_ptr(param, cmds[k%4].data_ptr(), nj)
        ar.collide_step(0.001)
        if mode=="late": ar.joints_set_param_batch_ptr(param, cmds[k%4].data_ptr(), nj)
        if mode in ("full","late","noh2d"):
            t=ar.observe_begin(sc["chassis"], obs[k&1].data_ptr(), jobs[k&1].data_ptr())
        else: t=None
        if pending is not None: ar.io_wait(pending)
        pending=t
    if pending is not None: ar.io_wait(pending)
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/K*1e3
for graph in (True, False):
    ar, sc = make(graph)
    cmds=[torch.empty((nj,W),dtype=torch.float32).pin_memory() for _ in range(4)]
    for i,c in enumerate(cmds): c[:]=torch.from_numpy(sc["vel"])[None,:]+0.001*i
    obs=[torch.empty((13,W),dtype=torch.float32).pin_memory() for _ in range(2)]
    jobs=[torch.empty((nj,4,W),dtype=torch.float32).pin_memory() for _ in range(2)]
    for mode in ("noh2d","full","late","h2d_io","full"):
        run(ar,sc,20,mode,cmds,obs,jobs); print("graph" if graph else "direct", mode, round(run(ar,sc,200,mode,cmds,obs,jobs),4), "ms/step", flush=True)
    del ar
