"""The reference's own test scene (test/reproducable_test.cpp) driven through the C++ facade
(WorldPhysics / Frame / Joint with the reference's interfaces) on the GPU, against the oracle's
restatement of the same scene (oracle/orc_mars.c).

The reference's criterion is determinism of the "%g" CSV (scripts/reproducable_test.sh); on top
of that the CSV is compared with the oracle's over a short horizon (fp32 device vs fp64 oracle:
the gait is chaotic, so the comparison horizon is bounded and the tolerance grows with it).
"""
import math
import os
import subprocess

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "mars_ode_physics_b200", "lib", "reproducable_test")


def _run(path, t_end, worlds=1, fast_step=1):
    assert os.path.exists(BIN), "facade test binary missing: run __graft_entry__.build()"
    subprocess.run([BIN, path, str(t_end), str(worlds), str(fast_step)], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    return np.loadtxt(path, delimiter=",", skiprows=1)


def test_reproducable_scene_deterministic(cuda_lib, tmp_path):
    a, b = str(tmp_path / "1.csv"), str(tmp_path / "2.csv")
    _run(a, 1.0)
    _run(b, 1.0)
    assert open(a).read() == open(b).read()  # scripts/reproducable_test.sh: `diff 1.csv i.csv`
    assert len(open(a).read().splitlines()) == 1001 + 1 or len(open(a).read().splitlines()) == 1000 + 1


def test_reproducable_scene_matches_oracle(cuda_lib, orc_lib, tmp_path):
    gpu = _run(str(tmp_path / "gpu.csv"), 0.5)
    ocsv = str(tmp_path / "orc.csv")
    rows = orc_lib.orc_mars_run_reproducable_test(ocsv.encode(), 0.5, orc.ORDER_ODE_ISLANDS, None)  # the facade's order
    ora = np.loadtxt(ocsv, delimiter=",", skiprows=1)
    assert rows == len(ora) == len(gpu)
    assert np.array_equal(gpu[:, 0], ora[:, 0])  # identical time column (same double accumulation)
    err = np.abs(gpu[:, 1:] - ora[:, 1:])
    print("facade vs oracle: max |dx| over 0.1 s", err[:100].max(), " over 0.5 s", err.max())
    assert err[:100].max() < 2e-4
    assert err.max() < 5e-3
    assert np.abs(ora[-1, 1:]).max() > 1e-3  # the robot actually moved


def test_reproducable_scene_with_fast_step_false(cuda_lib, orc_lib, tmp_path):
    """fast_step == false (the WorldPhysics constructor's default, reference src/WorldPhysics.cpp:82) selects
    dWorldStep (:367-370): through the facade that is the dense stepper (b2p_dense.cu), here on the reference's own
    test scene against the oracle's restatement of it (orc_lcp.c).  Deterministic like the quickstep run; the
    scene is over-constrained (84 rows, fixed joints, world cfm 1e-10), so the fp32 rows limit the agreement
    (tests/test_gpu_dense.py::test_dense_quadruped_with_friction states that sensitivity)."""
    a = _run(str(tmp_path / "d1.csv"), 0.2, 1, 0)
    b = _run(str(tmp_path / "d2.csv"), 0.2, 1, 0)
    assert np.array_equal(a, b)
    ocsv = str(tmp_path / "orc_dense.csv")
    rows = orc_lib.orc_mars_run_reproducable_test_stepper(ocsv.encode(), 0.2, orc.ORDER_ODE_ISLANDS, 1, None)
    ora = np.loadtxt(ocsv, delimiter=",", skiprows=1)
    quick = _run(str(tmp_path / "q.csv"), 0.2, 1, 1)
    assert rows == len(ora) == len(a)
    err = np.abs(a[:, 1:] - ora[:, 1:])
    print("facade fast_step=false vs oracle dWorldStep: max |dx| over 0.05 s", err[:50].max(), "over 0.2 s", err.max(),
          "| dense vs quickstep on the device", np.abs(a[:, 1:] - quick[:, 1:]).max())
    assert np.isfinite(a).all()
    assert err[:50].max() < 5e-4 and err.max() < 1e-2
    assert np.abs(a[:, 1:] - quick[:, 1:]).max() > 0.0  # it really is another stepper


def test_facade_replicas_agree(cuda_lib, tmp_path):
    """num_worlds > 1 replicates the template: world 0 of a 33-world arena = the one-world run."""
    one = _run(str(tmp_path / "one.csv"), 0.2, 1)
    many = _run(str(tmp_path / "many.csv"), 0.2, 33)
    assert np.array_equal(one, many)


def test_universal_and_ball_joints_through_the_loader(cuda_lib):
    """The joint types beyond the reference's registered set, created like a MARS host would (plugin loader ->
    createWorldInstance -> createFrame / createObject / createJoint with ConfigMaps): a two-link arm on a ball and
    a universal joint swings under gravity, the anchors hold, and the universal's first angle is observable."""
    exe = os.path.join(ROOT, "mars_ode_physics_b200", "lib", "joint_types_test")
    assert os.path.exists(exe), "facade test binary missing: run __graft_entry__.build()"
    out = subprocess.run([exe, "400"], check=True, capture_output=True, text=True).stdout.split()
    ball_err, len_a_err, len_b_err, angle1, moved = map(float, out)
    assert ball_err < 5e-3 and len_a_err < 5e-3 and len_b_err < 5e-3
    assert moved > 0.05 and math.isfinite(angle1)


def test_config_d_entirely_through_the_loader(cuda_lib, tmp_path):
    """BASELINE config D built and stepped through the plugin API only (facade_bench: WorldPhysicsLoader ->
    createFrame / createObject / createJoint with ConfigMaps, device collision from the objects' shapes, the
    heightfield entry point, ground_* surface defaults, stepTheWorld, batched commands and observations,
    ContactData read-back) gives the SAME worlds as the scene built through the C ABI by bench.build_rover_worlds:
    same bodies, geoms, joints and per-world randomisation, so the final states must agree bitwise."""
    import json
    import bench
    from mars_ode_physics_b200 import Arena
    exe = os.path.join(ROOT, "mars_ode_physics_b200", "lib", "facade_bench")
    assert os.path.exists(exe), "facade_bench missing: run __graft_entry__.build()"
    W, steps, warm, settle = 300, 40, 5, 200
    dump = str(tmp_path / "facade_state.bin")
    res = subprocess.run([exe, str(W), str(steps), str(warm), str(settle), dump], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:] + res.stdout[-500:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    print("facade_bench:", line)
    assert line["ok"] and line["worlds"] == W and line["e2e_facade"] > 0 and line["contacts_first_worlds"] >= 1
    assert line["launches"] == 4 * (settle + warm + steps)
    got = np.fromfile(dump, dtype=np.float32).reshape(5, 13, W)
    ar = Arena(W)
    sc = bench.build_rover_worlds(ar, W)
    from mars_ode_physics_b200 import capi
    ar.set_order_mode(capi.ORDER_ODE_ISLANDS)  # what the facade selects in initTheWorld
    param = capi.PARAM_VEL | capi.PARAM_GROUP2
    k_total = 0
    for _ in range(settle):
        ar.collide_step(0.001)
    for phase in (warm, steps):  # the bench rotates four command buffers (vel + 0.001 * (k % 4)) in both phases
        for k in range(phase):
            for j in sc["joints"]:
                ar.joint_set_param_batch(j, param, (sc["vel"] + np.float32(0.001) * np.float32(k % 4)).astype(np.float32))
            ar.collide_step(0.001)
            k_total += 1
    want = ar.state_download(5)
    d = np.abs(got.astype(np.float64) - want)
    print("facade vs C ABI rover worlds: max |diff|", d.max())
    assert np.array_equal(got, want)


def test_scene_construction_matches_the_reference_restatement(cuda_lib, orc_lib):
    """SURVEY 8 rows a11 / a12 through the plugin API (setup_test): every object type's mass pipeline -- sphere,
    capsule and cylinder with a rotated object pose, the three Inertial variants, and a composite frame whose
    centre of mass is re-centred twice (reference src/objects/*.cpp createMass, Frame.cpp:794-851) -- and the joint
    set-up: spring / damper constants mapped to stop ERP / CFM and motor fmax, jointCFM, the forced lo = hi = 0,
    hinge2's axis-2 values written over the axis-1 slots, a child-only (reversed) hinge, the friction coefficient
    (Joint.cpp:147-273 and the per-type createJoint).  Compared with oracle/orc_mars.c, which restates the same
    reference lines on the CPU."""
    import ctypes as C
    import json
    exe = os.path.join(ROOT, "mars_ode_physics_b200", "lib", "setup_test")
    assert os.path.exists(exe), "setup_test missing: run __graft_entry__.build()"
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    got = json.loads(res.stdout)

    L = orc_lib
    L.orc_mars_create.restype = C.c_void_p
    m = C.c_void_p(L.orc_mars_create())
    L.orc_mars_init_the_world(m)
    L.orc_mars_set_step_size(m, C.c_double(0.002))
    vec = lambda *v: (C.c_double * len(v))(*v)
    s30, c30 = math.sin(0.5 * 0.5235987755982988), math.cos(0.5 * 0.5235987755982988)

    def mass(kind, *a):
        mm = orc.Mass()
        {"sphere": lambda: L.orc_mass_set_sphere_total(C.byref(mm), C.c_double(a[0]), C.c_double(a[1])),
         "capsule": lambda: L.orc_mass_set_capsule_total(C.byref(mm), C.c_double(a[0]), 3, C.c_double(a[1]), C.c_double(a[2])),
         "cylinder": lambda: L.orc_mass_set_cylinder_total(C.byref(mm), C.c_double(a[0]), 3, C.c_double(a[1]), C.c_double(a[2])),
         "box": lambda: L.orc_mass_set_box_total(C.byref(mm), C.c_double(a[0]), C.c_double(a[1]), C.c_double(a[2]), C.c_double(a[3])),
         "box_density": lambda: L.orc_mass_set_box(C.byref(mm), C.c_double(a[0]), C.c_double(a[1]), C.c_double(a[2]), C.c_double(a[3])),
         }[kind]()
        return mm

    frames = {}

    def add(name, mm, pos, q=(1, 0, 0, 0)):
        if name not in frames:
            frames[name] = L.orc_mars_create_frame(m, C.c_double(1.1), C.c_double(1.1))
        L.orc_mars_frame_add_object(m, frames[name], C.byref(mm), vec(*pos), vec(*q))

    add("fs", mass("sphere", 2.0, 0.15), (0.1, 0.2, 0.3))
    add("fc", mass("capsule", 1.5, 0.05, 0.4), (0.0, 0.0, 1.0), (c30, s30, 0, 0))
    add("fy", mass("cylinder", 0.8, 0.1, 0.3), (0.5, 0.0, 0.5), (c30, 0, s30, 0))
    mi = orc.Mass()
    mi.mass = 1.2
    mi.I[:] = [0.02, 0.001, 0.0, 0.001, 0.03, 0.002, 0.0, 0.002, 0.025]
    add("fi", mi, (-0.3, 0.0, 0.2))
    add("fb", mass("box_density", 500.0, 0.2, 0.1, 0.05), (0.0, 0.4, 0.2))
    add("fsph", mass("sphere", 0.6, 0.07), (0.2, -0.4, 0.3))   # density 0 -> the mass branch
    add("comp", mass("box", 1.0, 0.4, 0.2, 0.1), (0.0, 0.0, 0.5))
    add("comp", mass("sphere", 0.3, 0.05), (0.2, 0.0, 0.6))
    add("comp", mass("cylinder", 0.5, 0.04, 0.2), (-0.1, 0.1, 0.45), (c30, s30, 0, 0))

    for name, fid in frames.items():
        mo, Io, po = C.c_double(), (C.c_double * 9)(), (C.c_double * 3)()
        L.orc_mars_frame_get_mass(m, fid, C.byref(mo), Io)
        L.orc_mars_frame_get_position(m, fid, po)
        g = got["frames"][name]
        assert abs(g["mass"] - mo.value) < 1e-12 * max(1.0, mo.value), (name, g["mass"], mo.value)
        assert np.allclose(g["I"], Io[:], rtol=1e-10, atol=1e-14), (name, g["I"], Io[:])
        assert np.allclose(g["frame_pos"], po[:], atol=1e-7), (name, g["frame_pos"], po[:])   # fp32 body position
        bo = (C.c_double * 3)()
        L.orc_body_get_position(C.c_void_p(L.orc_mars_world(m)), L.orc_mars_frame_body(m, fid), bo)
        assert np.allclose(g["body_pos"], bo[:], atol=1e-6), (name, g["body_pos"], bo[:])
    # composite: the frame still reports its origin although the body sits at the common centre of mass
    assert np.allclose(got["frames"]["comp"]["frame_pos"], 0.0, atol=1e-7)
    assert abs(got["frames"]["comp"]["mass"] - 1.8) < 1e-12

    L.orc_mars_world.restype = C.c_void_p
    L.orc_joint_get_param.restype = C.c_double
    w = C.c_void_p(L.orc_mars_world(m))

    def joint(jtype, parent, child, anchor, damping=None, stiffness=None, jcfm=None, friction=None):
        cfg = orc.MarsJointCfg()
        cfg.type = jtype
        cfg.parent = frames.get(parent, -1)
        cfg.child = frames.get(child, -1)
        cfg.anchor[:] = anchor
        cfg.axis1[:] = (0, 1, 0)
        cfg.axis2[:] = (0, 0, 1)
        for key, val in (("spring_damping", damping), ("spring_stiffness", stiffness), ("joint_cfm", jcfm)):
            setattr(cfg, "has_" + key, int(val is not None))
            setattr(cfg, key, float(val or 0.0))
        cfg.has_friction = int(friction is not None)
        cfg.friction_coefficient = float(friction or 0.0)
        j = L.orc_mars_create_joint(m, C.byref(cfg))
        assert j >= 0
        return L.orc_mars_joint_ode_id(m, j)

    P = orc
    want = {"hinge_spring_damper": joint(1, "fs", "fc", (0.05, 0.1, 0.6), 0.5, 20.0, 1e-4, 0.02),
            "slider_spring": joint(3, "fc", "fy", (0, 0, 0), None, 15.0),
            "hinge2_damper": joint(2, "comp", "fsph", (0.1, -0.2, 0.4), 0.3),
            "fixed": joint(6, "fi", "fb", (0, 0, 0)),
            "hinge_child_only": joint(1, "nonexistent", "fy", (0.5, 0.0, 0.7))}
    params = {"fmax": 5, "vel": 2, "lo": 0, "hi": 1, "stop_cfm": 10, "stop_erp": 9, "cfm": 8, "fmax2": 5 | 0x100,
              "lo2": 0 | 0x100}
    for name, jid in want.items():
        g = got["joints"][name]
        for key, pid in params.items():
            if name == "fixed" and key not in ("cfm",):
                continue
            wv = L.orc_joint_get_param(w, jid, pid)
            # (VEL / FMAX are per-world commands: the arena keeps them as fp32)
            tol = 1e-6 if key in ("fmax", "vel", "fmax2") else 1e-12
            assert (g[key] == wv) or abs(g[key] - wv) <= tol * max(1.0, abs(wv)), (name, key, g[key], wv)
        if name != "fixed" and name != "slider_spring":
            a = (C.c_double * 3)()
            L.orc_joint_get_anchor(w, jid, a)
            assert np.allclose(g["anchor"], a[:], atol=1e-6), (name, g["anchor"], a[:])
        if name != "fixed":
            ax = (C.c_double * 3)()
            L.orc_joint_get_axis1(w, jid, ax)
            assert np.allclose(g["axis1"], ax[:], atol=1e-6), (name, g["axis1"], ax[:])
    assert got["joints"]["hinge_spring_damper"]["friction"] == 0.02 and got["joints"]["fixed"]["friction"] == -0.1
    # the documented quirks made it through: damping as motor (fmax = damping, vel = 0), spring as stop with the
    # forced lo = hi = 0, hinge2 zeroing the axis-1 stop parameters
    assert got["joints"]["hinge_spring_damper"]["fmax"] == 0.5 and got["joints"]["hinge_spring_damper"]["lo"] == 0.0
    assert abs(got["joints"]["hinge2_damper"]["fmax2"] - 0.3) < 1e-7
    L.orc_mars_destroy(m)
