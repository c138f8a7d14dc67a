import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mars_ode_physics_b200 import Arena, capi
W = 65536
ar = Arena(W)
ar.set_stream(torch.cuda.current_stream().cuda_stream)
sc = bench.build_rover_worlds(ar, W)
for _ in range(600):
    ar.collide(); ar.step(0.001)
ar.sync()
for reorder in (capi.REORDER_NONE, capi.REORDER_RANDOMLY):
    ar.set_reorder_method(reorder)
    for iters in (1, 20):
        ar.set_quickstep_num_iterations(iters)
        for _ in range(5):
            ar.collide(); ar.step(0.001)
        ar.debug_kernel_times(True)
        for _ in range(30):
            ar.collide(); ar.step(0.001)
        kt = ar.debug_kernel_times(False)
        print("reorder", reorder, "iters", iters, "asm %.4f sor %.4f int %.4f" % (kt[0], kt[1], kt[2]), flush=True)
