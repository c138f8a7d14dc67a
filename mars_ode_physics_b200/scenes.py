"""Scene builders for the benchmark / parity configurations of BASELINE.json.

Every builder talks to a *backend* object that exposes the method names of
:class:`mars_ode_physics_b200.Arena` (body_create, joint_create, ...), so that the same scene
can be instantiated on the device arena and -- in tests only -- on the CPU oracle.
Only numpy is used here; nothing from ``oracle/`` is imported.

World parameters follow the reference defaults: gravity (0,0,-9.81), cfm 1e-10, erp 0.1
(src/WorldPhysics.cpp:83-85), max angular speed 10, contact max correcting velocity 5 (:95-96),
RNG seed 11211 (:169), 20 SOR iterations with w = 1.3 (ODE defaults, never overridden).
"""
import math

import numpy as np

from . import capi

CONTACT_MODE = capi.CONTACT_SOFT_ERP | capi.CONTACT_SOFT_CFM | capi.CONTACT_APPROX1


def splitmix64(x):
    """Deterministic per-world hash (seeds are `tag xor world_id`)."""
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def uniform01(tag, world, k=0):
    return (splitmix64((tag ^ world) + 0x1000003 * k) >> 11) / float(1 << 53)


def uniform01_array(tag, num_worlds, k=0):
    """Vectorised uniform01 over worlds 0..num_worlds-1 (uint64 arithmetic wraps like C)."""
    with np.errstate(over="ignore"):
        x = (np.uint64(tag) ^ np.arange(num_worlds, dtype=np.uint64)) + np.uint64(0x1000003 * k)
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def inertia_box(m, lx, ly, lz):
    return np.diag([m / 12.0 * (ly * ly + lz * lz), m / 12.0 * (lx * lx + lz * lz), m / 12.0 * (lx * lx + ly * ly)])


def inertia_sphere(m, r):
    return np.eye(3) * (0.4 * m * r * r)


def inertia_cylinder(m, r, length, axis=2):
    ia = m * (0.25 * r * r + length * length / 12.0)
    ib = 0.5 * m * r * r
    d = [ia, ia, ia]
    d[axis] = ib
    return np.diag(d)


def quat_axis_angle(axis, angle):
    axis = np.asarray(axis, dtype=float)
    axis = axis / np.linalg.norm(axis)
    s = math.sin(0.5 * angle)
    return np.array([math.cos(0.5 * angle), axis[0] * s, axis[1] * s, axis[2] * s])


def setup_world(be, iters=20, sor_w=1.3, seed=11211, cfm=1e-10, erp=0.1, gravity=(0.0, 0.0, -9.81)):
    be.set_gravity(*gravity)
    be.set_cfm(cfm)
    be.set_erp(erp)
    be.set_max_angular_speed(10.0)
    be.set_contact_max_correcting_vel(5.0)
    be.set_quickstep_num_iterations(iters)
    be.set_quickstep_w(sor_w)
    be.rand_set_seed(seed)


def add_body(be, mass, inertia, pos, quat=(1, 0, 0, 0)):
    b = be.body_create()
    be.body_set_mass(b, mass, inertia)
    be.body_set_position(b, pos)
    be.body_set_quaternion(b, quat)
    return b


def add_hinge(be, b1, b2, anchor, axis, fmax=None):
    j = be.joint_create(capi.JOINT_HINGE, b1, b2)
    be.joint_set_anchor(j, anchor)
    be.joint_set_axis1(j, axis)
    if fmax is not None:
        be.joint_set_param(j, capi.PARAM_FMAX, fmax)
    return j


# ---------------------------------------------------------------------------------------------
# config C: 12-hinge legged robot (13 bodies), laid out with the positions the reference test
# intended (test/reproducable_test.cpp:87-121): torso box, per leg upper/lower link + foot.
def quadruped(be, fmax=9.0, max_contacts=4):
    be.set_max_contacts(max_contacts)
    setup_world(be)
    sc = {"bodies": [], "joints": [], "feet": [], "foot_radius": 0.02, "name": "quadruped"}
    torso = add_body(be, 1.0, inertia_box(1.0, 0.4, 0.2, 0.05), (0.0, 0.0, 0.465))
    sc["torso"] = torso
    sc["bodies"].append(torso)
    hip, knee, ankle = [], [], []
    for sx in (1.0, -1.0):
        for sy in (1.0, -1.0):
            x, y = sx * 0.185, sy * 0.085
            upper = add_body(be, 0.2, inertia_box(0.2, 0.03, 0.03, 0.2), (x, y, 0.34))
            lower = add_body(be, 0.2, inertia_box(0.2, 0.03, 0.03, 0.2), (x, y, 0.14))
            foot = add_body(be, 0.05, inertia_sphere(0.05, 0.02), (x, y, 0.02))
            sc["bodies"] += [upper, lower, foot]
            sc["feet"].append(foot)
            hip.append(add_hinge(be, torso, upper, (x, y, 0.44), (0, 1, 0), fmax))
            knee.append(add_hinge(be, upper, lower, (x, y, 0.24), (0, 1, 0), fmax))
            ankle.append(add_hinge(be, lower, foot, (x, y, 0.04), (0, 1, 0), fmax))
    sc["hip"], sc["knee"], sc["ankle"] = hip, knee, ankle
    sc["joints"] = hip + knee + ankle
    return sc


def quadruped_targets(time, phase=0.0):
    """Sine gait of the reference test loop (test/reproducable_test.cpp:286-295)."""
    offs = (0.0, 1.14, 0.0, 1.14)
    hip = [0.5 * math.sin(phase + o + time * 4 * math.pi) - 0.3 for o in offs]
    knee = [-0.5 * math.sin(phase + o + 1 + time * 4 * math.pi) + 0.3 for o in offs]
    return hip, knee


def quadruped_control(be, sc, time, world, phase=0.0):
    """P-controller vel = 20*(target - position); hinge sign convention of Joint.cpp:362,402."""
    hip, knee = quadruped_targets(time, phase)
    for i in range(4):
        pos = -be.joint_get_angle1(sc["hip"][i], world)
        be.joint_set_param(sc["hip"][i], capi.PARAM_VEL, -((hip[i] - pos) * 20), world)
        pos = -be.joint_get_angle1(sc["knee"][i], world)
        be.joint_set_param(sc["knee"][i], capi.PARAM_VEL, -((knee[i] - pos) * 20), world)
        pos = -be.joint_get_angle1(sc["ankle"][i], world)
        be.joint_set_param(sc["ankle"][i], capi.PARAM_VEL, -((0.0 - pos) * 20), world)


def sphere_plane_contacts(be, bodies, radius, world, mu=0.8, soft_erp=0.2, soft_cfm=1e-5):
    """Host-side stand-in for narrowphase: sphere of `radius` at each body vs the plane z=0."""
    n = 0
    for b in bodies:
        p = be.body_get_position(b, world)
        depth = radius - p[2]
        if depth >= 0.0:
            be.contact_add(world, b, -1, pos=(p[0], p[1], p[2] - radius), normal=(0, 0, 1), depth=depth,
                           mode=CONTACT_MODE, mu=mu, mu2=mu, soft_erp=soft_erp, soft_cfm=soft_cfm)
            n += 1
    return n


# ---------------------------------------------------------------------------------------------
# config D: wheeled rover, chassis box + 4 wheels on hinge2 joints (axis1 = z steer, axis2 = y drive)
def rover(be, max_contacts=8, wheel_fmax=5.0, steer_lock=True, with_geoms=False, surface=None, wheel_shape="sphere"):
    """Config D rover: chassis box + 4 wheels on hinge2 joints (axis1 = z steer, locked by stops; axis2 = y drive).
    wheel_shape "sphere" (r 0.1, the bring-up variant SURVEY 8(d) allows) or "cylinder" (r 0.1, width 0.05, axis y)."""
    be.set_max_contacts(max_contacts)
    setup_world(be)
    sc = {"bodies": [], "joints": [], "wheels": [], "wheel_radius": 0.1, "name": "rover"}
    z0 = 0.25
    chassis = add_body(be, 10.0, inertia_box(10.0, 0.6, 0.4, 0.15), (0.0, 0.0, z0))
    sc["chassis"] = chassis
    sc["bodies"].append(chassis)
    for sx in (1.0, -1.0):
        for sy in (1.0, -1.0):
            x, y, z = sx * 0.25, sy * 0.25, z0 - 0.1
            wheel = add_body(be, 1.0, inertia_sphere(1.0, 0.1) if wheel_shape == "sphere"
                             else inertia_cylinder(1.0, 0.1, 0.05, axis=1), (x, y, z))
            sc["bodies"].append(wheel)
            sc["wheels"].append(wheel)
            j = be.joint_create(capi.JOINT_HINGE2, chassis, wheel)
            be.joint_set_anchor(j, (x, y, z))
            be.joint_set_axis1(j, (0, 0, 1))
            be.joint_set_axis2(j, (0, 1, 0))
            if steer_lock:
                be.joint_set_param(j, capi.PARAM_LO_STOP, 0.0)
                be.joint_set_param(j, capi.PARAM_HI_STOP, 0.0)
            be.joint_set_param(j, capi.PARAM_FMAX | capi.PARAM_GROUP2, wheel_fmax)
            be.joint_set_param(j, capi.PARAM_VEL | capi.PARAM_GROUP2, 4.0)
            be.joint_set_param(j, capi.PARAM_SUSPENSION_ERP, 0.4)
            be.joint_set_param(j, capi.PARAM_SUSPENSION_CFM, 0.01)
            sc["joints"].append(j)
    if with_geoms:
        s = surface or {"mu": 1.0, "mu2": 1.0, "soft_erp": 0.2, "soft_cfm": 1e-5}
        sc["geoms"] = [be.geom_create(capi.GEOM_BOX, chassis, (0.6, 0.4, 0.15), surface=s)]
        for wbody in sc["wheels"]:
            if wheel_shape == "sphere":
                sc["geoms"].append(be.geom_create(capi.GEOM_SPHERE, wbody, (0.1,), surface=s))
            else:  # the cylinder's axis (local z) turned onto the wheel axle (y)
                sc["geoms"].append(be.geom_create(capi.GEOM_CYLINDER, wbody, (0.1, 0.05),
                                                  lquat=quat_axis_angle((1, 0, 0), math.pi / 2), surface=s))
    return sc


def terrain_heights(nx=256, ny=256, dx=0.1, amplitude=0.1, seed=0xD):
    """Shared heightfield of config D: sum of 4 seeded sinusoids, amplitude 0.1 m."""
    xs = (np.arange(nx) - nx / 2) * dx
    ys = (np.arange(ny) - ny / 2) * dx
    X, Y = np.meshgrid(xs, ys)
    h = np.zeros_like(X)
    for k in range(4):
        kx = 0.5 + 2.5 * uniform01(seed, 100 + k, 1)
        ky = 0.5 + 2.5 * uniform01(seed, 100 + k, 2)
        ph = 2 * math.pi * uniform01(seed, 100 + k, 3)
        h += np.sin(kx * X + ky * Y + ph)
    h *= amplitude / 4.0
    return h.astype(np.float32), float(xs[0]), float(ys[0])


# ---------------------------------------------------------------------------------------------
# config B: 10-box stack on a plane
def box_stack(be, n=10, size=0.2, mass=1.0, max_contacts=48, with_geoms=True, surface=None):
    be.set_max_contacts(max_contacts)
    setup_world(be)
    sc = {"bodies": [], "joints": [], "name": "box_stack", "size": size}
    s = surface or {"mu": 0.8, "mu2": 0.8, "soft_erp": 0.2, "soft_cfm": 1e-5}
    if with_geoms:
        sc["plane"] = be.geom_create(capi.GEOM_PLANE, -1, (0, 0, 1, 0), surface=s)
    sc["geoms"] = []
    for i in range(n):
        z = 0.5 * size + (size + 0.005) * i
        b = add_body(be, mass, inertia_box(mass, size, size, size), (0.0, 0.0, z))
        sc["bodies"].append(b)
        if with_geoms:
            sc["geoms"].append(be.geom_create(capi.GEOM_BOX, b, (size, size, size), surface=s))
    return sc


def jitter_box_stack(be, sc, num_worlds, tag=0xB200):
    """Per-world jitter of config B: x,y ~ U(-0.01,0.01), yaw ~ U(-0.1,0.1)."""
    for w in range(num_worlds):
        for i, b in enumerate(sc["bodies"]):
            x = -0.01 + 0.02 * uniform01(tag, w, 3 * i + 0)
            y = -0.01 + 0.02 * uniform01(tag, w, 3 * i + 1)
            yaw = -0.1 + 0.2 * uniform01(tag, w, 3 * i + 2)
            z = 0.5 * sc["size"] + (sc["size"] + 0.005) * i
            be.body_set_position(b, (x, y, z), w)
            be.body_set_quaternion(b, quat_axis_angle((0, 0, 1), yaw), w)


# ---------------------------------------------------------------------------------------------
# small unit scenes for the per-joint-type parity tests
def pendulum_chain(be, n=3, joint_type=capi.JOINT_HINGE):
    setup_world(be)
    be.set_max_contacts(0)
    sc = {"bodies": [], "joints": [], "name": "pendulum_chain"}
    prev = -1
    for i in range(n):
        b = add_body(be, 0.5, inertia_box(0.5, 0.3, 0.05, 0.05), (0.15 + 0.3 * i, 0.0, 1.0))
        j = add_hinge(be, prev, b, (0.3 * i, 0.0, 1.0), (0, 1, 0))
        sc["bodies"].append(b)
        sc["joints"].append(j)
        prev = b
    return sc


def ball_universal_arm(be):
    """A two-link arm hanging from the world: link 0 on a ball joint (dJointCreateBall), link 1 on a universal
    joint (dJointCreateUniversal) whose first axis is motor-driven and whose second axis has stops."""
    setup_world(be)
    be.set_max_contacts(0)
    a = add_body(be, 0.8, inertia_box(0.8, 0.3, 0.06, 0.06), (0.15, 0.0, 1.0))
    b = add_body(be, 0.5, inertia_box(0.5, 0.3, 0.05, 0.05), (0.45, 0.0, 1.0))
    jb = be.joint_create(capi.JOINT_BALL, a, -1)
    be.joint_set_anchor(jb, (0.0, 0.0, 1.0))
    ju = be.joint_create(capi.JOINT_UNIVERSAL, a, b)
    be.joint_set_anchor(ju, (0.3, 0.0, 1.0))
    be.joint_set_axis1(ju, (0, 1, 0))
    be.joint_set_axis2(ju, (0, 0, 1))
    be.joint_set_param(ju, capi.PARAM_FMAX, 0.4)
    be.joint_set_param(ju, capi.PARAM_VEL, 0.8)
    be.joint_set_param(ju, capi.PARAM_LO_STOP | capi.PARAM_GROUP2, -0.2)
    be.joint_set_param(ju, capi.PARAM_HI_STOP | capi.PARAM_GROUP2, 0.25)
    return {"bodies": [a, b], "joints": [jb, ju], "ball": jb, "universal": ju, "name": "ball_universal_arm"}
