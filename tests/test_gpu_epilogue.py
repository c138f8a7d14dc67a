"""GPU parity of the round-2 device paths against the CPU oracle, through the C ABI:

* contact row variants Frame::addContact can produce (reference src/objects/Frame.cpp:1038-1079): rolling
  friction rows (rho, rho2, rhoN), Mu2 (mu != mu2), mu == 0, and the joint-connected rejection of
  WorldPhysics::createContact (src/WorldPhysics.cpp:458-468);
* what stepTheWorld does after the ODE call, for EVERY world of a batch (src/WorldPhysics.cpp:372-397):
  Joint::update, applyJointFriction, Frame::update (damping, accelerations), computeContactForce;
* dJointAddHingeTorque / dJointAddSliderForce, dParamBounce;
* the LCP residual of the device solve (north_star: "constraint residual after N iterations within a bound");
* host-side plumbing added with them: graph launch, one-world accessors that move one column, joint geometry
  setters that see the live pose, phantom threads of the Tensor-Memory SOR kernel.
"""
import numpy as np
import pytest

import orc
from mars_ode_physics_b200 import Arena, capi, scenes

pytestmark = pytest.mark.gpu

H = 0.001
MODE = scenes.CONTACT_MODE


def _state_err(a, b, nb):
    sa = a.state_download(nb).astype(np.float64)
    sb = np.asarray(b.state_download(nb), dtype=np.float64)
    d = np.abs(sa - sb)
    return {"pos": d[:, 0:3].max(), "quat": d[:, 3:7].max(), "lvel": d[:, 7:10].max(), "avel": d[:, 10:13].max()}


def _ball_on_plane(be, W, rolling):
    scenes.setup_world(be)
    be.set_max_contacts(2)
    if rolling:
        be.enable_rolling_friction(True)
    # 0.1 mm into the plane: a start exactly on it (depth 0) would be decided by fp32 rounding of z
    b = scenes.add_body(be, 1.0, scenes.inertia_sphere(1.0, 0.1), (0.0, 0.0, 0.0999))
    for w in range(W):
        be.body_set_linear_vel(b, (0.5 + 0.01 * w, -0.2, 0.0), w)
        be.body_set_angular_vel(b, (1.0, 0.5 * (w % 5), 2.0), w)
    return b


@pytest.mark.parametrize("variant", ["rolling", "rolling_mu2", "mu2_first_only", "mu2_second_only", "frictionless"])
def test_contact_row_variants(cuda_lib, variant):
    """One sphere per world sliding / rolling / spinning on the plane z = 0 with host-injected contacts in the
    mode-bit combinations Frame::addContact produces.  Row counts exact; states, contact feedback (rebuilt on
    the device from the normal and dPlaneSpace tangents) and the per-body contact force against the oracle."""
    W = 40
    kw = {"rolling": dict(mode=MODE | capi.CONTACT_ROLLING, mu=0.6, mu2=0.6, rho=0.02, rho2=0.02, rhoN=0.01, rows=6),
          "rolling_mu2": dict(mode=MODE | capi.CONTACT_ROLLING | capi.CONTACT_MU2, mu=0.6, mu2=0.3, rho=0.02,
                              rho2=0.05, rhoN=0.0, rows=5),
          "mu2_first_only": dict(mode=MODE | capi.CONTACT_MU2, mu=0.5, mu2=0.0, rows=2),
          "mu2_second_only": dict(mode=MODE | capi.CONTACT_MU2, mu=0.0, mu2=0.7, rows=2),
          "frictionless": dict(mode=MODE, mu=0.0, mu2=0.0, rows=1)}[variant]
    rows = kw.pop("rows")
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        b = _ball_on_plane(be, W, "rolling" in variant)
    worst_f = 0.0
    for step in range(80):
        for be in (ar, oc):
            be.contacts_clear()
            for w in range(W):
                p = be.body_get_position(b, w)
                depth = 0.1 - p[2]
                if depth >= 0:
                    be.contact_add(w, b, -1, pos=(p[0], p[1], p[2] - 0.1), normal=(0, 0, 1), depth=depth,
                                   soft_erp=0.2, soft_cfm=1e-5, **kw)
        for w in (0, 7, W - 1):
            assert ar.contact_count(w) == oc.contact_count(w)
        ar.step(H)
        oc.step(H)
        for w in (0, 7, W - 1):
            assert ar.last_num_rows(w) == oc.last_num_rows(w)
            if oc.contact_count(w):
                assert oc.last_num_rows(w) == rows, (variant, oc.last_num_rows(w))
                fa, fo = ar.contact_get_feedback(w, 0), oc.contact_get_feedback(w, 0)
                worst_f = max(worst_f, np.abs(fa - fo).max())
                assert np.allclose(ar.body_get_contact_force(b, w), fa, atol=1e-6)
                assert np.allclose(oc.body_get_contact_force(b, w), fo, atol=1e-12)
    assert sum(oc.contact_count(w) for w in range(W)) == W  # the balls stayed on the plane
    e = _state_err(ar, oc, 1)
    print(variant, "errors", e, "worst contact feedback error", worst_f)
    assert e["pos"] < 2e-5 and e["quat"] < 2e-5 and e["lvel"] < 2e-4 and e["avel"] < 2e-3, e
    assert worst_f < 2e-3, worst_f  # forces of ~10 N


def test_contact_between_joint_connected_bodies_is_rejected(cuda_lib):
    """WorldPhysics::createContact (reference :458-468): dAreConnectedExcluding(b1, b2, contact) -> no joint.
    b2p_contact_add reports B2P_EREJECTED (the wrapper returns -1) and adds nothing; contacts with the
    environment and between unconnected bodies are accepted."""
    W = 3
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        sc = scenes.pendulum_chain(be, 3)
        be.set_max_contacts(4)
    b = sc["bodies"]
    c = dict(pos=(0.3, 0, 1), normal=(0, 0, 1), depth=0.001, mode=MODE, mu=0.5, mu2=0.5, soft_erp=0.2, soft_cfm=1e-5)
    for be in (ar, oc):
        assert be.contact_add(1, b[0], b[1], **c) == -1
        assert be.contact_add(1, b[1], b[0], **c) == -1
        assert be.contact_count(1) == 0
        assert be.contact_add(1, b[0], b[2], **c) == 0       # not directly connected
        assert be.contact_add(1, b[1], -1, **c) == 1         # environment
        assert be.contact_count(1) == 2 and be.contact_count(0) == 0
    lib = cuda_lib
    d = capi.ContactDesc()
    assert lib.b2p_contact_add(ar.h, 1, b[0], b[1], d) == capi.E_REJECTED
    ar.step(H)
    oc.step(H)
    assert ar.last_num_rows(1) == oc.last_num_rows(1) == 15 + 6
    assert _state_err(ar, oc, 3)["pos"] < 1e-6


def test_joint_add_torque_and_bounce(cuda_lib):
    """dJointAddHingeTorque / dJointAddSliderForce (Joint::setTorque, reference src/joints/Joint.cpp:1121-1146)
    per world, and a hinge stop with dParamBounce, against the oracle."""
    W = 33
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        scenes.setup_world(be)
        be.set_max_contacts(0)
        a = scenes.add_body(be, 0.5, scenes.inertia_box(0.5, 0.3, 0.05, 0.05), (0.15, 0.0, 1.0))
        s = scenes.add_body(be, 0.7, scenes.inertia_box(0.7, 0.1, 0.1, 0.1), (0.0, 1.0, 1.0))
        c = scenes.add_body(be, 0.4, scenes.inertia_box(0.4, 0.3, 0.05, 0.05), (0.15, 2.0, 1.0))
        jh = scenes.add_hinge(be, -1, a, (0.0, 0.0, 1.0), (0, 1, 0))       # reversed hinge (body 1 = environment)
        js = be.joint_create(capi.JOINT_SLIDER, s, -1)
        be.joint_set_axis1(js, (1, 0, 0))
        jb = scenes.add_hinge(be, c, -1, (0.0, 2.0, 1.0), (0, 1, 0))       # swings into bouncing stops
        be.joint_set_param(jb, capi.PARAM_LO_STOP, -0.15)
        be.joint_set_param(jb, capi.PARAM_HI_STOP, 0.1)
        be.joint_set_param(jb, capi.PARAM_BOUNCE, 0.6)
    for step in range(150):
        for be in (ar, oc):
            for w in range(0, W, 4):
                be.joint_add_torque(jh, 0.3 + 0.02 * w, w)
                be.joint_add_torque(js, -0.5 + 0.05 * w, w)
        ar.step(H)
        oc.step(H)
    e = _state_err(ar, oc, 3)
    print("add_torque / bounce errors", e)
    assert e["pos"] < 5e-5 and e["quat"] < 1e-4 and e["lvel"] < 2e-3 and e["avel"] < 1e-2, e
    for w in (0, 4, W - 1):
        assert abs(ar.joint_get_angle1(jb, w) - oc.joint_get_angle1(jb, w)) < 1e-4
        assert ar.joint_get_num_rows(jb, w) == oc.joint_get_num_rows(jb, w)
    # the stop was reached and the bounce reversed the swing in at least one world
    assert max(oc.joint_get_num_rows(jb, w) for w in range(W)) == 6 or abs(oc.joint_get_angle1(jb, 0)) < 0.2


def test_post_step_bookkeeping_for_every_world(cuda_lib):
    """Joint::update / applyJointFriction / Frame::update / computeContactForce run on the device for every world
    (the reference does them on the host for its one world): damping < 1 on all bodies, joint friction > 0 on
    the hips, W > 1, compared world by world -- including worlds other than 0 -- with the oracle."""
    W = 24
    ar, oc = Arena(W), orc.OracleBatch(W)
    scs = []
    for be in (ar, oc):
        sc = scenes.quadruped(be)
        scs.append(sc)
        for b in sc["bodies"]:
            be.body_set_damping(b, 0.995, 0.99)
        for j in sc["hip"]:
            be.joint_set_friction_coefficient(j, 0.05)
    sa, so = scs
    nb, nj = len(sa["bodies"]), len(sa["joints"])
    t = 0.0
    for step in range(60):
        for be, sc in ((ar, sa), (oc, so)):
            be.contacts_clear()
            for w in range(W):
                scenes.quadruped_control(be, sc, t, w, 0.4 * w)
                scenes.sphere_plane_contacts(be, sc["feet"], sc["foot_radius"], w)
        ar.step(H)
        oc.step(H)
        t += H
    e = _state_err(ar, oc, nb)
    print("damped + joint friction quadruped errors", e)
    assert e["pos"] < 2e-4 and e["quat"] < 1e-3 and e["lvel"] < 2e-2 and e["avel"] < 0.2, e
    # damping and friction did act (an undamped run differs)
    ref = Arena(W)
    sr = scenes.quadruped(ref)
    t = 0.0
    for step in range(60):
        ref.contacts_clear()
        for w in range(W):
            scenes.quadruped_control(ref, sr, t, w, 0.4 * w)
            scenes.sphere_plane_contacts(ref, sr["feet"], sr["foot_radius"], w)
        ref.step(H)
        t += H
    diff = np.abs(ref.state_download(nb) - ar.state_download(nb))[:, 7:, :].max(axis=(0, 1))
    assert (diff > 1e-3).all(), diff  # every world, not only world 0
    # telemetry arrays, all worlds at once
    pa, po = ar.body_post_download(nb).astype(np.float64), oc.body_post_download(nb)
    ja, jo = ar.joint_update_download(nj).astype(np.float64), oc.joint_update_download(nj)
    assert np.abs(pa[:, 0:6] - po[:, 0:6]).max() < 0.2                      # last velocities
    acc_scale = np.abs(po[:, 6:12]).max()
    assert np.abs(pa[:, 6:12] - po[:, 6:12]).max() < 2e-2 * acc_scale + 50.0, acc_scale  # finite differences / 1 ms
    assert np.allclose(pa[:, 12:15], po[:, 12:15], rtol=5e-2, atol=5e-2)     # contact force per body
    assert np.abs(po[:, 12:15]).max() > 0.5
    assert np.allclose(ja, jo, rtol=5e-2, atol=5e-2), np.abs(ja - jo).max()
    # one-world accessors agree with the batched download
    w = W - 3
    assert np.allclose(ar.joint_get_update(sa["hip"][1], w), ja[sa["hip"][1], :, w], atol=1e-6)
    la, aa = ar.body_get_acceleration(sa["torso"], w)
    assert np.allclose(np.concatenate([la, aa]), pa[sa["torso"], 6:12, w], atol=1e-3)


def test_reversed_joint_update_matches_reference_quirk(cuda_lib):
    """A joint whose first body is the environment takes the reference's "else if (body2)" arm, which reads
    feedback.f2 / t2 -- zero for a one-body joint -- and therefore produces 0/0 (Joint.cpp:944-984).  The device
    reproduces that: same NaN pattern as the oracle, finite motor torque."""
    W = 5
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        sc = scenes.pendulum_chain(be, 2)
        be.joint_set_param(sc["joints"][0], capi.PARAM_FMAX, 0.2)
        be.joint_set_param(sc["joints"][0], capi.PARAM_VEL, 0.5)
    for _ in range(5):
        ar.step(H)
        oc.step(H)
    ja, jo = ar.joint_update_download(2).astype(np.float64), oc.joint_update_download(2)
    assert np.array_equal(np.isnan(ja), np.isnan(jo))
    assert np.isnan(jo[0, 1:]).all() and not np.isnan(jo[1]).any()
    assert np.allclose(ja[1], jo[1], rtol=2e-3, atol=2e-3)
    assert np.allclose(ja[0, 0], jo[0, 0], rtol=2e-3, atol=1e-4) and np.abs(jo[0, 0]).max() > 1e-3


def test_device_lcp_residual(cuda_lib):
    """The residual of the device's own solve (computed on the device from its rows, multipliers and constraint
    forces) equals the oracle's for the same step, and stays under the bound stated for 20 iterations on the
    stiff 12-hinge scene (tests/test_oracle.py::test_residual_decreases_with_iterations: 2e4); more iterations
    lower it."""
    W = 16
    worst = {}
    for iters in (20, 200):
        ar, oc = Arena(W), orc.OracleBatch(W)
        scs = []
        for be in (ar, oc):
            scs.append(scenes.quadruped(be))
            be.set_quickstep_num_iterations(iters)
        t, hi = 0.0, 0.0
        for step in range(30):
            for be, sc in zip((ar, oc), scs):
                be.contacts_clear()
                for w in range(W):
                    scenes.quadruped_control(be, sc, t, w, 0.3 * w)
                    scenes.sphere_plane_contacts(be, sc["feet"], sc["foot_radius"], w)
            ar.step(H)
            oc.step(H)
            t += H
            ra = ar.lcp_residual().astype(np.float64)
            ro = np.array([oc.last_residual(w) for w in range(W)])
            assert np.allclose(ra, ro, rtol=5e-2, atol=2.0), (step, np.abs(ra - ro).max(), ro.max())
            hi = max(hi, ra.max())
        worst[iters] = hi
    print("device LCP residual (max over 30 steps x 16 worlds):", worst)
    assert worst[20] < 2e4
    assert worst[200] < 0.5 * worst[20]


def test_rover_residual_bound_at_full_size(cuda_lib):
    """BASELINE config D: residual of the device solve over all 65536 worlds after the rovers have landed."""
    import bench
    W = 65536
    ar = Arena(W)
    bench.build_rover_worlds(ar, W)
    for _ in range(300):
        ar.collide_step(H)
    r = ar.lcp_residual()
    print("rover residual: max", r.max(), "median", np.median(r), "99.9%", np.quantile(r, 0.999))
    assert np.isfinite(r).all()
    # observed on B200: median 1.4e-6, 99.9 % quantile 62, max 181 (worlds in hard contact transitions)
    assert np.median(r) < 1e-3 and np.quantile(r, 0.999) < 600.0 and r.max() < 2e3
    assert ar.contacts_dropped() == 0


def test_graph_launch_is_bitwise_equal_to_direct_launch(cuda_lib):
    """collide + step replayed from a CUDA graph (default) produce bitwise the states of direct launches, also
    across parameter changes (step size) that force a re-capture."""
    import bench
    outs = []
    for graph in (True, False):
        ar = Arena(300)
        sc = bench.build_rover_worlds(ar, 300)
        ar.set_graph_launch(graph)
        for k in range(120):
            ar.collide_step(H if k < 80 else 0.5 * H)
        for k in range(20):
            ar.collide()
            ar.step(H)
        outs.append((ar.state_download(len(sc["bodies"])).copy(), [ar.rand_get_seed(w) for w in (0, 150, 299)],
                     ar.kernel_launch_count()))
        ar.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert outs[0][2] == outs[1][2] == 4 * 140


def test_one_world_accessors_after_a_step(cuda_lib):
    """After a step the device copy is the current one: one-world setters / getters move one column, not the
    batch, and must not disturb the other worlds (Frame::setLinearVelocity, addForce, Joint::setVelocity ...)."""
    W = 70
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        sc = scenes.pendulum_chain(be, 2)
        be.joint_set_param(sc["joints"][1], capi.PARAM_FMAX, 0.3)
    b = sc["bodies"]
    for step in range(30):
        for be in (ar, oc):
            if step == 10:
                be.body_set_linear_vel(b[1], (0.1, 0.2, 0.3), 5)
                be.body_set_angular_vel(b[0], (0.0, 1.0, 0.0), 69)
                be.body_set_quaternion(b[1], (0.99, 0.0, 0.1, 0.0), 33)
                be.body_set_position(b[1], (0.5, 0.0, 1.02), 33)
            if step >= 10:
                be.body_add_force(b[1], (0.0, 0.0, 2.0), 7)
                be.body_add_force(b[1], (0.0, 0.0, 1.0), 7)       # accumulates
                be.body_add_torque(b[0], (0.0, 0.1, 0.0), 64)
                be.joint_set_param(sc["joints"][1], capi.PARAM_VEL, 0.5 if step % 2 else -0.5, 12)
        if step == 12:
            assert np.allclose(ar.body_get_force(b[1], 7), (0, 0, 3.0)) and np.allclose(ar.body_get_force(b[1], 8), 0)
            assert ar.joint_get_param(sc["joints"][1], capi.PARAM_VEL, 12) == -0.5
        ar.step(H)
        oc.step(H)
    assert np.allclose(ar.body_get_force(b[1], 7), 0)  # the step zeroed the accumulators
    e = _state_err(ar, oc, 2)
    assert e["pos"] < 2e-5 and e["quat"] < 5e-5 and e["lvel"] < 1e-3 and e["avel"] < 5e-3, e


def test_joint_geometry_setters_use_the_live_pose(cuda_lib):
    """dJointSetHingeAnchor / dJointSetFixed read the bodies' CURRENT poses (ODE) -- what Joint::reattacheJoint
    relies on (reference src/joints/Joint.cpp:782-825).  After the worlds have moved, re-setting the anchor at
    its current world position must leave the joint satisfied (no jump), as on the oracle."""
    W = 2
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        sc = scenes.pendulum_chain(be, 2)
        f = scenes.add_body(be, 0.3, scenes.inertia_box(0.3, 0.1, 0.1, 0.1), (0.75, 0.0, 1.0))
        jf = be.joint_create(capi.JOINT_FIXED, sc["bodies"][1], f)
        be.joint_set_fixed(jf)
    for _ in range(120):
        ar.step(H)
        oc.step(H)
    for be in (ar, oc):
        p = be.joint_get_anchor(sc["joints"][1], 0)
        be.joint_set_anchor(sc["joints"][1], p)      # reattacheJoint of a hinge
        be.joint_set_fixed(jf)                       # reattacheJoint of a fixed joint
    for _ in range(40):
        ar.step(H)
        oc.step(H)
    e = _state_err(ar, oc, 3)
    print("reattach errors", e)
    assert e["pos"] < 1e-4 and e["quat"] < 2e-4 and e["lvel"] < 5e-3, e
    # no jump: the anchor points of the re-attached hinge still coincide
    assert np.abs(ar.joint_get_anchor(sc["joints"][1], 0) - ar.joint_get_anchor2(sc["joints"][1], 0)).max() < 1e-3


def test_tmem_kernel_phantom_threads_leave_other_ctas_alone(cuda_lib):
    """A CTA of the Tensor-Memory SOR kernel that takes fewer worlds than it has threads runs phantom threads;
    with worlds that have NO rows mixed among worlds that have many (so that phantoms share warps with long
    worlds and finish late) they must not touch the worlds of the next CTA.  Compared with the paired-lane
    kernel, bitwise, including the per-world RNG state."""
    W = 148 * 2 + 77  # several CTAs with worlds_per_cta < 256
    outs = []
    for kernel in (capi.SOR_TMEM, capi.SOR_PAIRED):
        ar = Arena(W)
        sc = scenes.pendulum_chain(ar, 3)
        ar.set_max_contacts(2)
        ar.set_sor_kernel(kernel)
        c = dict(normal=(0, 0, 1), depth=0.001, mode=MODE, mu=0.5, mu2=0.5, soft_erp=0.2, soft_cfm=1e-5)
        for step in range(12):
            ar.contacts_clear()
            for w in range(0, W, 3):
                ar.contact_add(w, sc["bodies"][2], -1, pos=(0.9, 0, 1), **c)
            ar.step(H)
        outs.append((ar.state_download(3).copy(), [ar.rand_get_seed(w) for w in range(W)],
                     [ar.last_num_rows(w) for w in range(0, W, 7)]))
        assert ar.sor_kernel_used() == kernel
        ar.close()
    assert outs[0][2] == outs[1][2]
    assert outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][0], outs[1][0])
    # the same with worlds that have no rows at all next to worlds that do (free bodies + one contact)
    outs = []
    for kernel in (capi.SOR_TMEM, capi.SOR_PAIRED):
        ar = Arena(W)
        scenes.setup_world(ar)
        ar.set_max_contacts(4)
        ar.set_sor_kernel(kernel)
        bs = [scenes.add_body(ar, 1.0, scenes.inertia_sphere(1.0, 0.1), (i, 0, 0.1)) for i in range(3)]
        for step in range(20):
            ar.contacts_clear()
            for w in range(W):
                if (w * 7919) % 5 < 2:  # 40 % of the worlds have contacts, scattered
                    for b in bs[: 1 + w % 3]:
                        p = ar.body_get_position(b, w)
                        ar.contact_add(w, b, -1, pos=(p[0], p[1], 0.0), normal=(0, 0, 1), depth=max(0.1 - p[2], 0.0),
                                       mode=MODE, mu=0.5, mu2=0.5, soft_erp=0.2, soft_cfm=1e-5)
            ar.step(H)
        outs.append((ar.state_download(3).copy(), [ar.rand_get_seed(w) for w in range(W)]))
        ar.close()
    assert outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][0], outs[1][0])


def _multi_island_scene(be, W):
    """Three separate pendulum chains (three islands with rows), two free bodies (islands without rows) and a ball
    on the plane with a host-injected contact: the row blocks and the order in which the world's LCG is consumed
    differ between ODE's island order and the one-block-per-world order."""
    scenes.setup_world(be)
    be.set_max_contacts(4)
    bodies, joints = [], []
    for k, n in enumerate((2, 3, 1)):
        prev = -1
        for i in range(n):
            b = scenes.add_body(be, 0.5, scenes.inertia_box(0.5, 0.3, 0.05, 0.05), (0.15 + 0.3 * i, 1.0 * k, 1.0))
            j = scenes.add_hinge(be, prev, b, (0.3 * i, 1.0 * k, 1.0), (0, 1, 0), 0.2 if i == 0 else None)
            bodies.append(b)
            joints.append(j)
            prev = b
        if k == 0:  # a free body created between the chains
            bodies.append(scenes.add_body(be, 1.0, scenes.inertia_sphere(1.0, 0.1), (5.0, 5.0, 5.0)))
    ball = scenes.add_body(be, 1.0, scenes.inertia_sphere(1.0, 0.1), (0.0, -2.0, 0.0999))
    bodies.append(ball)
    bodies.append(scenes.add_body(be, 1.0, scenes.inertia_sphere(1.0, 0.1), (-5.0, 5.0, 5.0)))
    for w in range(W):
        be.body_set_angular_vel(bodies[1], (0, 0.05 * w, 0.0), w)
        be.body_set_linear_vel(ball, (0.3, 0.01 * w, 0.0), w)
    return bodies, joints, ball


@pytest.mark.parametrize("kernel", [capi.SOR_TMEM, capi.SOR_PAIRED])
def test_ode_island_order_matches_oracle(cuda_lib, kernel):
    """B2P_ORDER_ODE_ISLANDS: rows in ODE's dxProcessIslands order (bodies from the most recently created one,
    depth-first through the joint lists, the step's contacts first), one row block per island, per-island
    reshuffles that consume the world's LCG island by island -- against the oracle's ORDER_ODE_ISLANDS.  The RNG
    state after every step is an exact check of the block sizes and of the draw order."""
    W = 37
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        bodies, joints, ball = _multi_island_scene(be, W)
        be.set_order_mode(capi.ORDER_ODE_ISLANDS)
    ar.set_sor_kernel(kernel)
    nb = len(bodies)
    for step in range(60):
        for be in (ar, oc):
            be.contacts_clear()
            for w in range(W):
                p = be.body_get_position(ball, w)
                if w % 3 != 2 and 0.1 - p[2] >= 0:  # every third world has no contact: one island less
                    be.contact_add(w, ball, -1, pos=(p[0], p[1], p[2] - 0.1), normal=(0, 0, 1), depth=0.1 - p[2],
                                   mode=MODE, mu=0.5, mu2=0.5, soft_erp=0.2, soft_cfm=1e-5)
        ar.step(H)
        oc.step(H)
        for w in (0, 1, 2, 17, W - 1):
            assert ar.last_num_rows(w) == oc.last_num_rows(w), (step, w)
            assert ar.rand_get_seed(w) == oc.rand_get_seed(w), (step, w)
    assert ar.sor_kernel_used() == kernel
    e = _state_err(ar, oc, nb)
    print("island order, multi-island scene:", e)
    assert e["pos"] < 2e-5 and e["quat"] < 5e-5 and e["lvel"] < 1e-3 and e["avel"] < 5e-3, e
    fa, fo = ar.joint_get_feedback(joints[3], 5), oc.joint_get_feedback(joints[3], 5)
    assert np.allclose(fa, fo, rtol=2e-3, atol=2e-3), (fa, fo)
    # and the island order does differ from the world-block order (different sweep order, different RNG use)
    ref = orc.OracleBatch(1)
    _multi_island_scene(ref, 1)
    ref.step(H)
    isl = orc.OracleBatch(1)
    _multi_island_scene(isl, 1)
    isl.set_order_mode(capi.ORDER_ODE_ISLANDS)
    isl.step(H)
    assert np.abs(ref.state_download(nb) - isl.state_download(nb)).max() > 0


def test_ode_island_order_with_device_collision(cuda_lib):
    """The island order on scenes whose islands are made by the device contact stream: config D rovers (one
    island; ODE lists the step's contacts before the joints) and the config B box stack (ten islands in free fall
    that merge into one when the stack lands), against the oracle's ORDER_ODE_ISLANDS; and the two SOR kernels
    agree bitwise in this mode too."""
    import bench
    W = 40
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        sc = bench.build_rover_worlds(be, W)
        be.set_order_mode(capi.ORDER_ODE_ISLANDS)
    nb = len(sc["bodies"])
    for step in range(220):
        ar.collide_step(H)
        oc.collide_step(H)
        if step % 20 == 19:
            for w in (0, 13, W - 1):
                if ar.contact_count(w) == oc.contact_count(w):
                    assert ar.last_num_rows(w) == oc.last_num_rows(w), (step, w)
    e = _state_err(ar, oc, nb)
    print("island order, rovers on the heightfield, 220 steps:", e)
    assert e["pos"] < 5e-4 and e["quat"] < 1e-3, e
    outs = []
    for kernel in (capi.SOR_TMEM, capi.SOR_PAIRED):
        a2 = Arena(64)
        s2 = scenes.box_stack(a2, n=10)
        scenes.jitter_box_stack(a2, s2, 64)
        a2.set_order_mode(capi.ORDER_ODE_ISLANDS)
        a2.set_sor_kernel(kernel)
        for step in range(60):
            a2.collide_step(H)
        outs.append((a2.state_download(10).copy(), [a2.rand_get_seed(w) for w in range(64)]))
        a2.close()
    assert outs[0][1] == outs[1][1] and np.array_equal(outs[0][0], outs[1][0])
    # box stack against the oracle through the landing: identical inputs (state and RNG) re-imposed before every
    # compared step, row counts and RNG states exact whenever the contact counts agree
    Wb = 16
    ab, ob = Arena(Wb), orc.OracleBatch(Wb)
    for be in (ab, ob):
        sb = scenes.box_stack(be, n=10)
        scenes.jitter_box_stack(be, sb, Wb)
        be.set_order_mode(capi.ORDER_ODE_ISLANDS)
    checked = 0
    for step in range(60):
        st32 = ob.state_download(10).astype(np.float32)
        ab.state_upload(st32)
        for w in range(Wb):
            for b in range(10):
                ob.body_set_state(b, st32[b, :, w].astype(np.float64), w)
        ab.rand_set_seed(1000 + step)
        ob.rand_set_seed(1000 + step)
        ab.collide()
        ob.collide()
        same = [ab.contact_count(w) == ob.contact_count(w) for w in range(Wb)]
        ab.step(H)
        ob.step(H)
        for w in range(Wb):
            if same[w]:
                assert ab.last_num_rows(w) == ob.last_num_rows(w), (step, w)
                assert ab.rand_get_seed(w) == ob.rand_get_seed(w), (step, w)
                checked += 1
    assert checked > 50 * Wb


@pytest.mark.gpu
def test_integrate_shared_memory_stash_changes_nothing(cuda_lib):
    """The integrate kernel keeps the contact-force sums and Joint::update's inputs in shared memory when they fit
    (nb*3 + nj*7 <= 128 columns); larger scenes read them back from global memory.  Both paths on the same scene
    (tune bit 4 forces the second one), the SOR experiment switches along the way: bitwise equal."""
    W = 96
    outs = []
    for tune in (0, 4, 1 | 2 | 4):
        ar = Arena(W)
        ar.set_tune(tune)
        sc = scenes.quadruped(ar)
        for b in sc["bodies"]:
            ar.body_set_damping(b, 0.995, 0.99)
        for j in sc["hip"]:
            ar.joint_set_friction_coefficient(j, 0.05)
        nb, nj = len(sc["bodies"]), len(sc["joints"])
        t = 0.0
        for step in range(40):
            ar.contacts_clear()
            for w in range(W):
                scenes.quadruped_control(ar, sc, t, w, 0.4 * w)
                scenes.sphere_plane_contacts(ar, sc["feet"], sc["foot_radius"], w)
            ar.step(H)
            t += H
        outs.append((ar.state_download(nb), ar.body_post_download(nb), ar.joint_update_download(nj)))
    for o in outs[1:]:
        for x, y in zip(outs[0], o):
            assert np.array_equal(x, y)


@pytest.mark.gpu
def test_device_arrays_as_cuda_array_interface_and_dlpack(cuda_lib):
    """The batched observation / telemetry arrays are visible to other GPU code without a copy (CUDA array interface,
    DLPack): what torch sees through them equals the C ABI's batched downloads."""
    torch = pytest.importorskip("torch")
    W = 40
    ar = Arena(W)
    sc = scenes.quadruped(ar)
    nb, nj = len(sc["bodies"]), len(sc["joints"])
    t = 0.0
    for step in range(10):
        ar.contacts_clear()
        for w in range(W):
            scenes.quadruped_control(ar, sc, t, w, 0.3 * w)
            scenes.sphere_plane_contacts(ar, sc["feet"], sc["foot_radius"], w)
        ar.step(H)
        t += H
    ar.sync()
    obs = torch.as_tensor(ar.device_array("joint_obs", nj), device="cuda")
    assert obs.shape == (nj, 4, ar.padded_worlds()) and obs.dtype == torch.float32
    assert np.array_equal(obs[:, :, :W].cpu().numpy(), ar.joint_obs_download(nj))
    st = torch.from_dlpack(ar.device_array("state", nb))
    assert np.array_equal(st[:, :, :W].cpu().numpy(), ar.state_download(nb))
    post = torch.from_dlpack(ar.device_array("body_post", nb))
    assert np.array_equal(post[:, :, :W].cpu().numpy(), ar.body_post_download(nb))
    cnt = torch.as_tensor(ar.device_array("contact_count"), device="cuda")
    assert cnt.dtype == torch.int32 and [int(c) for c in cnt[:W].cpu()] == [ar.contact_count(w) for w in range(W)]
