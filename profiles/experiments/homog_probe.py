# SOR kernel time for homogeneous batches: every world with the same number of rows (flat terrain: 4 wheel contacts;
# no terrain contact: 24 rows) against the mixed batch of the bench
import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch, bench
from mars_ode_physics_b200 import Arena, scenes
W=65536
def run(amp, lift):
    orig = scenes.terrain_heights
    scenes.terrain_heights = lambda *a, **k: (lambda h,x,y: ((h*amp).astype(np.float32), x, y))(*orig(*a, **k))
    ar=Arena(W); ar.set_stream(torch.cuda.current_stream().cuda_stream)
    sc=bench.build_rover_worlds(ar,W); ar.set_order_mode(1)
    scenes.terrain_heights = orig
    if lift:
        nb=len(sc["bodies"]); st=np.array(ar.state_download(nb)); st[:,2,:]+=50.0; ar.state_upload(st)
        ar.set_gravity(0,0,0) if hasattr(ar,"set_gravity") else None
    for _ in range(300 if not lift else 5): ar.collide_step(0.001)
    ar.sync()
    ar.debug_kernel_times(True)
    for _ in range(30): ar.collide_step(0.001)
    kt=ar.debug_kernel_times(False)
    nr=np.array([ar.last_num_rows(w) for w in range(0,W,64)])
    print(f"amp {amp} lift {lift}: rows mean {nr.mean():.2f} min {nr.min()} max {nr.max()}  assemble {kt[0]:.4f} sor {kt[1]:.4f} integrate {kt[2]:.4f} ms", flush=True)
run(1.0, False)
run(0.0, False)
run(1.0, True)
