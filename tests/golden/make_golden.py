"""Regenerates the regression fixtures under tests/golden/ from the oracle.

These are NOT reference outputs (the reference stores no golden vectors and libode is absent:
"parity unpinned"); they pin the oracle against accidental change.
  reproducable_test_oracle_f64.csv : first 1001 rows (1 s) of the reference test scene
  rover_terrain_oracle_f64.npz     : state of 4 rover worlds after 200 collide+step cycles
  arm_ball_universal_oracle_f64.npz: ball + universal arm (scenes.ball_universal_arm), 6 worlds, 300 steps: state,
                                     the universal's angles / rates, per-step row counts
  sampled_colliders_oracle_f64.npz : contact records of the sample-point colliders (capsule / cylinder pairs) on
                                     fixed seeded poses
  dense_stepper_oracle_f64.npz     : dWorldStep (fast_step == false): rover on the heightfield (4 worlds, 320 cycles:
                                     state, per-step contact and row counts) and friction-less quadruped (3 worlds,
                                     60 steps: state)
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

# ---- ball + universal arm
W = 6
oc = orc.OracleBatch(W)
sc = scenes.ball_universal_arm(oc)
for w in range(W):
    oc.body_set_angular_vel(sc["bodies"][1], (0.0, 0.3 * w, 3.0 - 0.5 * w), w)
rows = []
for _ in range(300):
    oc.step(0.001)
    rows.append([oc.last_num_rows(w) for w in range(W)])
ju = sc["universal"]
np.savez_compressed(os.path.join(HERE, "arm_ball_universal_oracle_f64.npz"), state=oc.state_download(2),
                    rows=np.array(rows, dtype=np.int32),
                    angles=np.array([[oc.joint_get_angle1(ju, w), oc.joint_get_angle1_rate(ju, w),
                                      oc.joint_get_angle2(ju, w), oc.joint_get_angle2_rate(ju, w)] for w in range(W)]),
                    seeds=np.array([oc.rand_get_seed(w) for w in range(W)], dtype=np.uint32))

# ---- sample-point colliders: contact records on seeded poses
from mars_ode_physics_b200 import capi  # noqa: E402
pairs = [(capi.GEOM_CAPSULE, (0.05, 0.3), capi.GEOM_BOX, (0.3, 0.2, 0.25)),
         (capi.GEOM_SPHERE, (0.1,), capi.GEOM_CYLINDER, (0.1, 0.25)),
         (capi.GEOM_CAPSULE, (0.05, 0.3), capi.GEOM_CYLINDER, (0.1, 0.25)),
         (capi.GEOM_CYLINDER, (0.1, 0.25), capi.GEOM_BOX, (0.3, 0.2, 0.25)),
         (capi.GEOM_CYLINDER, (0.1, 0.25), capi.GEOM_CYLINDER, (0.12, 0.2))]
surf = {"mu": 0.8, "mu2": 0.8, "soft_erp": 0.2, "soft_cfm": 1e-5}
recs, poses = [], []
for pi, (ca, da, cb, db) in enumerate(pairs):
    W = 24
    rng = np.random.default_rng(900 + pi)
    oc = orc.OracleBatch(W)
    scenes.setup_world(oc)
    oc.set_max_contacts(8)
    a = scenes.add_body(oc, 1.0, np.eye(3) * 0.01, (0, 0, 0))
    oc.geom_create(ca, a, da, surface=surf)
    b = scenes.add_body(oc, 1.0, np.eye(3) * 0.01, (0, 0, 0))
    oc.geom_create(cb, b, db, surface=surf)
    pose = np.zeros((W, 2, 7))
    for w in range(W):
        for k, body in enumerate((a, b)):
            p = rng.uniform(-0.12, 0.12, size=3)
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            pose[w, k, :3], pose[w, k, 3:] = p, q
            oc.body_set_position(body, p, w)
            oc.body_set_quaternion(body, q, w)
    oc.collide()
    rec = np.full((W, 8, 8), np.nan)
    for w in range(W):
        for c in range(oc.contact_count(w)):
            b1, b2, d = oc.contact_get(w, c)
            rec[w, c] = [b1, *d.pos[:], *d.normal[:], d.depth]
    recs.append(rec)
    poses.append(pose)
np.savez_compressed(os.path.join(HERE, "sampled_colliders_oracle_f64.npz"), records=np.array(recs), poses=np.array(poses))
print("golden fixtures written")

# ---- dWorldStep (dense stepper, orc_lcp.c): the rover on the heightfield, 4 worlds, 320 collide+step cycles from the
# drop (the wheels land after about 160), and the friction-less quadruped, 3 worlds, 60 steps
W = 4
oc = orc.OracleBatch(W)
sc = scenes.rover(oc, with_geoms=True)
oc.geom_create_heightfield(h, 0.1, x0, y0, surface={"mu": 1.0, "mu2": 1.0, "soft_erp": 0.2, "soft_cfm": 1e-5})
oc.set_stepper(1)
for w in range(W):
    for b in sc["bodies"]:
        p = oc.body_get_position(b, w)
        oc.body_set_position(b, (p[0] + 0.7 * w, p[1] - 0.4 * w, p[2] + 0.1), w)
    for j in sc["joints"]:
        oc.joint_set_param(j, scenes.capi.PARAM_VEL | scenes.capi.PARAM_GROUP2, 2.0 + w, w)
counts, rows = [], []
for _ in range(320):
    oc.collide()
    counts.append([oc.contact_count(w) for w in range(W)])
    oc.step(0.001)
    rows.append([oc.last_num_rows(w) for w in range(W)])
rover_state = oc.state_download(5)
W = 3
oc = orc.OracleBatch(W)
sc = scenes.quadruped(oc)
oc.set_stepper(1)
t = 0.0
for _ in range(60):
    oc.contacts_clear()
    for w in range(W):
        scenes.quadruped_control(oc, sc, t, w, 0.37 * w)
        scenes.sphere_plane_contacts(oc, sc["feet"], sc["foot_radius"], w, mu=0.0)
    oc.step(0.001)
    t += 0.001
np.savez_compressed(os.path.join(HERE, "dense_stepper_oracle_f64.npz"), rover_state=rover_state,
                    rover_counts=np.array(counts, dtype=np.int32), rover_rows=np.array(rows, dtype=np.int32),
                    quadruped_state=oc.state_download(13))
