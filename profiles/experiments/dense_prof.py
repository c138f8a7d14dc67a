import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from mars_ode_physics_b200 import Arena, capi
W = 65536
ar = Arena(W)
bench.build_rover_worlds(ar, W)
ar.set_order_mode(1)
for _ in range(400): ar.collide_step(0.001)
ar.set_stepper(1)
ar.set_graph_launch(False)
for _ in range(3): ar.collide_step(0.001)
ar.sync()
print("rows", np.mean([ar.last_num_rows(w) for w in range(0, W, 512)]))
