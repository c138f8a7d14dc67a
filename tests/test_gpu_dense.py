"""GPU parity of the dense stepper (dWorldStep, fast_step == false; src/WorldPhysics.cpp:367-370): the CUDA path
(assemble kernel -> b2p_dense_lcp_kernel -> integrate kernel, through the C ABI) against the oracle's restatement of
ODE's step.cpp / lcp.cpp (oracle/orc_lcp.c) on the same seeded inputs.  The device assembles its rows in fp32 and
solves the LCP in fp64; the oracle is fp64 throughout.  Integer work (row and contact counts) must match exactly."""
import numpy as np
import pytest

import orc
from mars_ode_physics_b200 import Arena, capi, scenes

pytestmark = pytest.mark.gpu

H = 0.001


def _err(ar, oc, nb):
    sa = ar.state_download(nb).astype(np.float64)
    so = oc.state_download(nb)
    d = np.abs(sa - so)
    return {"pos": d[:, 0:3].max(), "quat": d[:, 3:7].max(), "lvel": d[:, 7:10].max(), "avel": d[:, 10:13].max()}


def _run_dense_quadruped(be_a, be_o, W, steps, mu, order=None, exact_counts=True):
    sa, so = scenes.quadruped(be_a), scenes.quadruped(be_o)
    be_a.set_stepper(capi.STEPPER_DENSE)  # (the oracle's ORC_STEPPER_DENSE has the same value)
    be_o.set_stepper(capi.STEPPER_DENSE)
    if order is not None:
        be_a.set_order_mode(order)
    t = 0.0
    for step in range(steps):
        for be, sc in ((be_a, sa), (be_o, so)):
            be.contacts_clear()
            for w in range(W):
                scenes.quadruped_control(be, sc, t, w, 0.37 * w)
                scenes.sphere_plane_contacts(be, sc["feet"], sc["foot_radius"], w, mu=mu)
        same = [be_a.contact_count(w) == be_o.contact_count(w) for w in range(W)]
        assert all(same) or not exact_counts, (step, same)
        be_a.step(H)
        be_o.step(H)
        t += H
        for w in (0, W // 2, W - 1):
            assert be_a.last_num_rows(w) == be_o.last_num_rows(w) or not same[w], (step, w)
    return len(sa["bodies"])


@pytest.mark.parametrize("order", [capi.ORDER_WORLD_BLOCK, capi.ORDER_ODE_ISLANDS])
def test_dense_quadruped_parity_frictionless(cuda_lib, order):
    """12-hinge quadruped (config C topology: 60 unbounded hinge rows, 12 bounded motor rows, up to 4 one-sided
    normal rows -- 76 rows on 78 degrees of freedom), 33 worlds with different gait phases, 120 steps, device
    (rows fp32, LCP fp64) against the fp64 oracle.  Both solve the same well-posed LCP, so the trajectories agree
    to fp32 rounding; the device's multipliers satisfy its LCP to 1e-12; no world makes the solver give up.  The
    row order (world block / ODE islands) does not matter to the exact solver."""
    W = 33
    ar, oc = Arena(W), orc.OracleBatch(W)
    nb = _run_dense_quadruped(ar, oc, W, 120, 0.0, order)
    assert ar.stepper() == capi.STEPPER_DENSE
    assert ar.dense_failures() == 0 and all(oc.last_lcp_error(w) == 0 for w in range(W))
    res = float(ar.dense_residual().max())
    e = _err(ar, oc, nb)
    print("dense quadruped, friction-less:", e, "max scaled LCP residual", res)
    # observed on B200: pos 1.2e-6, quat 1.0e-6, lvel 4e-5, avel 3e-4; the bars are ~10x that
    assert res < 1e-9
    assert e["pos"] < 1e-5 and e["quat"] < 1e-5 and e["lvel"] < 5e-4 and e["avel"] < 3e-3, e


def test_dense_quadruped_with_friction(cuda_lib):
    """The same scene with friction (mu 0.8): 84 rows on 78 degrees of freedom -- a rank-deficient LCP matrix whose
    only regularisation is the world cfm of 1e-10 (soft cfm 1e-5 on the normal rows).  Its solution is sensitive to
    the rounding of the rows: the oracle built with fp32 rows (same solver, in double) deviates from the fp64 oracle
    by the same amount as the device does, which is what this test states -- the device stays within 3x of that
    sensitivity over 10 steps -- next to the exact things: contact and row counts, the device's LCP residual, no
    solver failure (a contact at depth ~ 0 may exist on one side only once the states differ by 1e-5)."""
    W = 8
    ar, oc = Arena(W), orc.OracleBatch(W)
    nb = _run_dense_quadruped(ar, oc, W, 10, 0.8, exact_counts=False)  # a foot at depth ~ 0 may be in or out
    o32, o64 = orc.OracleBatch(W, single=True), orc.OracleBatch(W)
    _run_dense_quadruped(o32, o64, W, 10, 0.8, exact_counts=False)  # the fp32-row oracle stands where the device stood
    d = np.abs(o32.state_download(nb).astype(np.float64) - o64.state_download(nb))
    sens = {"pos": d[:, :3].max(), "lvel": d[:, 7:10].max(), "avel": d[:, 10:13].max()}
    e = _err(ar, oc, nb)
    res = float(ar.dense_residual().max())
    print("dense quadruped with friction: device vs fp64 oracle", e, "| fp32-row oracle vs fp64 oracle", sens,
          "| residual", res)
    assert ar.dense_failures() == 0
    assert res < 1e-9
    assert e["pos"] < 3 * sens["pos"] + 1e-6 and e["lvel"] < 3 * sens["lvel"] + 1e-5 and e["avel"] < 3 * sens["avel"] + 1e-4, (e, sens)
    assert e["pos"] < 5e-4 and e["lvel"] < 0.2, e  # and in absolute terms


def test_dense_rover_on_heightfield(cuda_lib):
    """Config D scene with device collision (sphere wheels on the shared heightfield, friction-dependent rows): the
    dense stepper on the device against the oracle's, 320 steps from the drop (the wheels land after about 160); contact counts exact while the
    states agree, poses within tolerance."""
    import bench
    W = 48
    ar, oc = Arena(W), orc.OracleBatch(W)
    sa, so = bench.build_rover_worlds(ar, W), bench.build_rover_worlds(oc, W)
    ar.set_stepper(capi.STEPPER_DENSE)
    oc.set_stepper(1)
    nb = len(sa["bodies"])
    for step in range(320):
        ar.collide()
        oc.collide()
        if step % 40 == 0:
            ca = [ar.contact_count(w) for w in range(W)]
            co = [oc.contact_count(w) for w in range(W)]
            assert sum(1 for x, y in zip(ca, co) if x != y) <= 1, (step, ca, co)
        ar.step(H)
        oc.step(H)
    assert ar.dense_failures() == 0
    assert all(oc.last_lcp_error(w) == 0 for w in range(W))
    res = float(ar.dense_residual().max())
    e = _err(ar, oc, nb)
    rows = [ar.last_num_rows(w) for w in range(W)]
    print("dense rover:", e, "residual", res, "rows", rows)
    assert sum(1 for r in rows if r > 24) > W // 4  # wheels touching down: contact and friction rows present
    # observed on B200: pos 3.0e-5, quat 2.9e-5, lvel 1.0e-3, avel 2.4e-3 (the rover bounces: contacts come and go)
    assert res < 1e-9
    assert e["pos"] < 3e-4 and e["quat"] < 3e-4 and e["lvel"] < 1e-2 and e["avel"] < 3e-2, e


def test_dense_graph_launch_and_switching(cuda_lib):
    """The stepper can be switched between steps (the step graph is re-captured), and the dense step replayed from
    the CUDA graph equals the directly launched one bitwise."""
    import bench
    outs = []
    for graph in (True, False):
        ar = Arena(40)
        sc = bench.build_rover_worlds(ar, 40)
        ar.set_graph_launch(graph)
        for step in range(60):
            ar.set_stepper(capi.STEPPER_DENSE if (step // 10) % 2 else capi.STEPPER_QUICK)
            ar.collide()
            ar.step(H)
        outs.append(ar.state_download(len(sc["bodies"])).copy())
        ar.close()
    assert np.isfinite(outs[0]).all()
    assert np.array_equal(outs[0], outs[1])


def test_dense_warp_kernel_equals_thread_kernel(cuda_lib):
    """The two dense kernels (one warp per world in shared memory; one thread per world in global scratch, the
    fallback for scenes beyond ~160 row slots) run the same pivoting sequence and differ only in the order of their
    floating-point sums: after 200 steps of the rover scene (contacts, friction, index removals) and 40 steps of the
    quadruped (84 rows, three position-vector elements per lane) the states agree to 1e-5 and both solve their
    LCPs."""
    import bench
    outs = []
    for tune in (0, 8):
        ar = Arena(70)
        sc = bench.build_rover_worlds(ar, 70)
        ar.set_tune(tune)
        ar.set_stepper(capi.STEPPER_DENSE)
        for _ in range(200):
            ar.collide()
            ar.step(H)
        assert ar.dense_failures() == 0
        assert float(ar.dense_residual().max()) < 1e-9
        outs.append(ar.state_download(len(sc["bodies"])).astype(np.float64))
        ar.close()
    d = np.abs(outs[0] - outs[1])
    print("rover: warp kernel vs thread kernel", d[:, :7].max(), d[:, 7:].max())
    assert d[:, :7].max() < 1e-5 and d[:, 7:].max() < 1e-3
    outs = []
    for tune in (0, 8):
        ar = Arena(9)
        sc = scenes.quadruped(ar)
        ar.set_tune(tune)
        ar.set_stepper(capi.STEPPER_DENSE)
        t = 0.0
        for _ in range(40):
            ar.contacts_clear()
            for w in range(9):
                scenes.quadruped_control(ar, sc, t, w, 0.37 * w)
                scenes.sphere_plane_contacts(ar, sc["feet"], sc["foot_radius"], w, mu=0.0)
            ar.step(H)
            t += H
        assert ar.dense_failures() == 0
        outs.append(ar.state_download(len(sc["bodies"])).astype(np.float64))
        ar.close()
    d = np.abs(outs[0] - outs[1])
    print("quadruped: warp kernel vs thread kernel", d[:, :7].max(), d[:, 7:].max())
    assert d[:, :7].max() < 1e-6 and d[:, 7:].max() < 1e-4


def test_dense_against_committed_fixture(cuda_lib):
    """The committed oracle fixture of the dense stepper (tests/golden/dense_stepper_oracle_f64.npz, generator
    tests/golden/make_golden.py): the device reproduces the friction-less quadruped to fp32 rounding and the rover
    on the heightfield -- 320 free-running steps with contacts coming and going -- within 1e-4, with the same contact
    and row counts in at least 97 % of the (step, world) pairs (a wheel at depth ~ 0 may be in or out once the states
    differ by 1e-5)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dense_stepper_oracle_f64.npz"))
    W = 4
    ar = Arena(W)
    sc = scenes.rover(ar, with_geoms=True)
    h, x0, y0 = scenes.terrain_heights()
    ar.geom_create_heightfield(h, 0.1, x0, y0, surface={"mu": 1.0, "mu2": 1.0, "soft_erp": 0.2, "soft_cfm": 1e-5})
    ar.set_stepper(capi.STEPPER_DENSE)
    for w in range(W):
        for b in sc["bodies"]:
            p = ar.body_get_position(b, w)
            ar.body_set_position(b, (p[0] + 0.7 * w, p[1] - 0.4 * w, p[2] + 0.1), w)
        for j in sc["joints"]:
            ar.joint_set_param(j, capi.PARAM_VEL | capi.PARAM_GROUP2, 2.0 + w, w)
    same = total = 0
    for k in range(320):
        ar.collide()
        cc = [ar.contact_count(w) for w in range(W)]
        ar.step(H)
        rr = [ar.last_num_rows(w) for w in range(W)]
        same += sum(int(a == b) for a, b in zip(cc, g["rover_counts"][k])) + sum(int(a == b) for a, b in zip(rr, g["rover_rows"][k]))
        total += 2 * W
    assert ar.dense_failures() == 0
    d = np.abs(ar.state_download(5).astype(np.float64) - g["rover_state"])
    print("dense rover vs fixture: pose", d[:, :7].max(), "velocity", d[:, 7:].max(), "equal counts", same, "of", total)
    assert same >= 0.97 * total
    # observed on B200: pose 9.1e-6, velocity 1.3e-3, all 2560 counts equal
    assert d[:, :7].max() < 1e-4 and d[:, 7:].max() < 1.5e-2
    W = 3
    ar = Arena(W)
    sc = scenes.quadruped(ar)
    ar.set_stepper(capi.STEPPER_DENSE)
    t = 0.0
    for _ in range(60):
        ar.contacts_clear()
        for w in range(W):
            scenes.quadruped_control(ar, sc, t, w, 0.37 * w)
            scenes.sphere_plane_contacts(ar, sc["feet"], sc["foot_radius"], w, mu=0.0)
        ar.step(H)
        t += H
    d = np.abs(ar.state_download(13).astype(np.float64) - g["quadruped_state"])
    print("dense quadruped vs fixture: pose", d[:, :7].max(), "velocity", d[:, 7:].max())
    assert d[:, :7].max() < 1e-5 and d[:, 7:].max() < 3e-3


def test_dense_edge_cases(cuda_lib):
    """Worlds without any row (free bodies: the dense stepper must leave the ballistic motion of quickstep untouched,
    bitwise), a batch in which only some worlds have rows, and a world count that leaves the last warp / CTA ragged."""
    outs = []
    for stepper in (capi.STEPPER_QUICK, capi.STEPPER_DENSE):
        W = 37
        ar = Arena(W)
        scenes.setup_world(ar)
        ar.set_max_contacts(4)
        for i in range(3):
            scenes.add_body(ar, 1.0 + i, scenes.inertia_box(1.0 + i, 0.3, 0.2, 0.1), (i, 0, 1))
        for w in range(W):
            for b in range(3):
                ar.body_set_linear_vel(b, (0.1 * w, -0.2 * b, 1.0), w)
                ar.body_set_angular_vel(b, (0.5 * b, 0.3 * w, -1.0), w)
        ar.set_stepper(stepper)
        for _ in range(50):
            ar.step(H)
        outs.append(ar.state_download(3).copy())
        ar.close()
    assert np.array_equal(outs[0], outs[1])
    # contacts in every third world only: a body resting on one frictional contact is held, the others fall freely
    W = 37
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        scenes.setup_world(be)
        be.set_max_contacts(4)
        scenes.add_body(be, 2.0, scenes.inertia_box(2.0, 0.2, 0.2, 0.2), (0, 0, 0.1))
        be.set_stepper(capi.STEPPER_DENSE)
    for _ in range(30):
        for be in (ar, oc):
            be.contacts_clear()
            for w in range(0, W, 3):
                be.contact_add(w, 0, -1, pos=(0, 0, 0.0), normal=(0, 0, 1), depth=0.001, mode=scenes.CONTACT_MODE,
                               mu=0.8, mu2=0.8, soft_erp=0.2, soft_cfm=1e-5)
        ar.step(H)
        oc.step(H)
    assert ar.dense_failures() == 0
    assert [ar.last_num_rows(w) for w in range(W)] == [3 if w % 3 == 0 else 0 for w in range(W)]
    e = _err(ar, oc, 1)
    print("dense, ragged batch:", e)
    assert e["pos"] < 1e-6 and e["lvel"] < 1e-5, e
