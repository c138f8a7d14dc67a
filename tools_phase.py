import sys, numpy as np
sys.path.insert(0, '.')
import bench
from mars_ode_physics_b200 import Arena, capi
W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for mode in ("random", "none"):
    ar = Arena(W)
    sc = bench.build_rover_worlds(ar, W)
    if mode == "none":
        ar.set_reorder_method(capi.REORDER_NONE)
    for _ in range(620):
        ar.collide(); ar.step(0.001)
    ar.debug_phase_cycles(True)
    for _ in range(50):
        ar.collide(); ar.step(0.001)
    c = ar.debug_phase_cycles(False)
    tot = c[:5].sum()
    print(mode, "cycles/warp-step: assembly %.0f  reshuffle+permute %.0f  sweeps %.0f  writeback %.0f  epilogue %.0f  total %.0f (%.3f ms @1.965GHz)" % (*c[:5], tot, tot / 1.965e6))
    ar.close()
