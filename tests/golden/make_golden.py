"""Regenerates the regression fixtures under tests/golden/ from the oracle.

These are NOT reference outputs (the reference stores no golden vectors and libode is absent:
"parity unpinned"); they pin the oracle against accidental change.
  reproducable_test_oracle_f64.csv : first 1001 rows (1 s) of the reference test scene
  rover_terrain_oracle_f64.npz     : state of 4 rover worlds after 200 collide+step cycles
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import orc  # noqa: E402
from mars_ode_physics_b200 import scenes  # noqa: E402

lib = orc.load(False)
tmp = os.path.join(HERE, "_tmp.csv")
lib.orc_mars_run_reproducable_test(tmp.encode(), 1.0, orc.ORDER_WORLD_BLOCK, None)
lines = open(tmp).read().splitlines()[:1002]
open(os.path.join(HERE, "reproducable_test_oracle_f64.csv"), "w").write("\n".join(lines) + "\n")
os.remove(tmp)

W = 4
oc = orc.OracleBatch(W)
sc = scenes.rover(oc, with_geoms=True)
h, x0, y0 = scenes.terrain_heights()
oc.geom_create_heightfield(h, 0.1, x0, y0, surface={"mu": 1.0, "mu2": 1.0, "soft_erp": 0.2, "soft_cfm": 1e-5})
for w in range(W):
    for b in sc["bodies"]:
        p = oc.body_get_position(b, w)
        oc.body_set_position(b, (p[0] + 0.7 * w, p[1] - 0.4 * w, p[2] + 0.1), w)
counts = []
for _ in range(200):
    oc.collide()
    counts.append([oc.contact_count(w) for w in range(W)])
    oc.step(0.001)
np.savez_compressed(os.path.join(HERE, "rover_terrain_oracle_f64.npz"), state=oc.state_download(5),
                    counts=np.array(counts, dtype=np.int32), seeds=np.array([oc.rand_get_seed(w) for w in range(W)], dtype=np.uint32))
print("golden fixtures written")
