import sys, os, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np, torch, bench
from mars_ode_physics_b200 import Arena, capi, scenes
H = 0.001
for name, builder, W, settle in (("B box stack", bench.build_box_stack_worlds, 256, 150), ("C quadruped", bench.build_quadruped_worlds, 2048, 100)):
    ar = Arena(W)
    builder(ar, W, world_offset=0)
    for _ in range(settle): ar.collide_step(H)
    ar.set_stepper(1)
    ar.sync(); t0 = time.perf_counter()
    for _ in range(50): ar.collide_step(H)
    ar.sync(); dt = time.perf_counter() - t0
    nb = 10 if name[0] == "B" else 13
    st = ar.state_download(nb)
    rows = [ar.last_num_rows(w) for w in range(0, W, max(1, W // 64))]
    print(name, "W", W, "ms/step", dt / 50 * 1e3, "failures", ar.dense_failures(), "of", W * 50, "finite", bool(np.isfinite(st).all()), "max residual", float(ar.dense_residual().max()), "rows mean/max", np.mean(rows), max(rows), "max |v|", float(np.abs(st[:, 7:10]).max()))
    ar.close()
