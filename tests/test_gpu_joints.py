"""GPU parity per joint type: hinge / slider with motors and
stops (limit rows, powered-into-stop force injection), fixed, hinge2 with both limit motors, joints
attached to the static environment in both argument orders (dJOINT_REVERSE)."""
import math

import numpy as np
import pytest

import orc
from mars_ode_physics_b200 import Arena, capi, scenes

pytestmark = pytest.mark.gpu
H = 0.001


def _pair(W, single=False, strict=False):
    return Arena(W, strict=strict), orc.OracleBatch(W, single=single)


def _compare(ar, oc, nb, tol_pos=2e-4, tol_vel=2e-2):
    sa = ar.state_download(nb).astype(np.float64)
    so = oc.state_download(nb)
    d = np.abs(sa - so)
    assert d[:, :7].max() < tol_pos, ("pose", d[:, :7].max())
    assert d[:, 7:].max() < tol_vel, ("vel", d[:, 7:].max())
    return d


def _step_both(ar, oc, n, W, joints=()):
    for s in range(n):
        ar.step(H)
        oc.step(H)
        if s % 25 == 0:
            for w in range(0, W, max(1, W // 4)):
                assert ar.last_num_rows(w) == oc.last_num_rows(w), (s, w)
                for j in joints:
                    assert ar.joint_get_num_rows(j, w) == oc.joint_get_num_rows(j, w), (s, w, j)


def test_hinge_stops_and_motor_into_stop(cuda_lib):
    W = 16
    ar, oc = _pair(W)
    for be in (ar, oc):
        scenes.setup_world(be, gravity=(0, 0, 0))  # the motor direction decides which stop is hit
        be.set_max_contacts(0)
        a = scenes.add_body(be, 1.0, scenes.inertia_box(1.0, 0.2, 0.2, 0.2), (0, 0, 1))
        b = scenes.add_body(be, 0.5, scenes.inertia_box(0.5, 0.4, 0.05, 0.05), (0.3, 0, 1))
        j0 = be.joint_create(capi.JOINT_FIXED, a, -1)
        be.joint_set_fixed(j0)
        j = scenes.add_hinge(be, a, b, (0.1, 0, 1), (0, 1, 0))
        be.joint_set_param(j, capi.PARAM_LO_STOP, -0.3)
        be.joint_set_param(j, capi.PARAM_HI_STOP, 0.4)
        be.joint_set_param(j, capi.PARAM_STOP_ERP, 0.3)
        be.joint_set_param(j, capi.PARAM_STOP_CFM, 1e-4)
        be.joint_set_param(j, capi.PARAM_FUDGE_FACTOR, 0.5)
        for w in range(W):  # motors of different strength / direction: some worlds are driven into a stop
            be.joint_set_param(j, capi.PARAM_FMAX, 0.2 * (w % 4), w)
            be.joint_set_param(j, capi.PARAM_VEL, -2.0 + 0.3 * w, w)
    _step_both(ar, oc, 600, W, joints=(1,))
    _compare(ar, oc, 2)
    angles = [oc.joint_get_angle1(1, w) for w in range(W)]
    assert min(angles) < -0.25 and max(angles) > 0.3  # both stops were reached
    for w in range(0, W, 3):
        assert abs(ar.joint_get_angle1(1, w) - oc.joint_get_angle1(1, w)) < 2e-4
        assert abs(ar.joint_get_angle1_rate(1, w) - oc.joint_get_angle1_rate(1, w)) < 2e-2
        fa, fo = ar.joint_get_feedback(1, w), oc.joint_get_feedback(1, w)
        assert np.allclose(fa, fo, rtol=2e-2, atol=2e-2), (w, fa, fo)


def test_slider_motor_stops_and_reverse_attach(cuda_lib):
    W = 12
    ar, oc = _pair(W)
    for be in (ar, oc):
        scenes.setup_world(be)
        be.set_max_contacts(0)
        a = scenes.add_body(be, 1.0, scenes.inertia_box(1.0, 0.2, 0.2, 0.2), (0, 0, 1), scenes.quat_axis_angle((0, 0, 1), 0.3))
        b = scenes.add_body(be, 0.5, scenes.inertia_box(0.5, 0.1, 0.1, 0.1), (0.3, 0.1, 1.1))
        c = scenes.add_body(be, 0.7, scenes.inertia_box(0.7, 0.1, 0.1, 0.1), (-0.5, 0, 1.0))
        js = be.joint_create(capi.JOINT_SLIDER, -1, a)  # reversed: environment first
        be.joint_set_axis1(js, (0, 0, 1))
        be.joint_set_param(js, capi.PARAM_LO_STOP, -0.05)
        be.joint_set_param(js, capi.PARAM_HI_STOP, 0.08)
        jb = be.joint_create(capi.JOINT_SLIDER, a, b)
        be.joint_set_axis1(jb, (1, 0.2, 0))
        be.joint_set_param(jb, capi.PARAM_FMAX, 3.0)
        jc = be.joint_create(capi.JOINT_SLIDER, c, -1)  # body first
        be.joint_set_axis1(jc, (0, 1, 1))
        be.joint_set_param(jc, capi.PARAM_LO_STOP, -0.02)
        be.joint_set_param(jc, capi.PARAM_HI_STOP, 0.02)
        be.joint_set_param(jc, capi.PARAM_FMAX, 20.0)
        for w in range(W):
            be.joint_set_param(jb, capi.PARAM_VEL, -0.5 + 0.1 * w, w)
            be.joint_set_param(jc, capi.PARAM_VEL, 0.3 if w % 2 else -0.3, w)
    _step_both(ar, oc, 400, W, joints=(0, 1, 2))
    _compare(ar, oc, 3)
    for w in (0, 5, W - 1):
        for j in range(3):
            assert abs(ar.joint_get_angle1(j, w) - oc.joint_get_angle1(j, w)) < 2e-4, (j, w)
            assert abs(ar.joint_get_angle1_rate(j, w) - oc.joint_get_angle1_rate(j, w)) < 2e-2, (j, w)
            assert np.allclose(ar.joint_get_axis1(j, w), oc.joint_get_axis1(j, w), atol=1e-4)


def test_hinge2_both_axes(cuda_lib):
    """Rover-like hinge2 with a free (motorised) steering axis in some worlds, locked in others."""
    W = 8
    ar, oc = _pair(W)
    for be in (ar, oc):
        sc = scenes.rover(be, steer_lock=False)
        for jn, j in enumerate(sc["joints"]):
            be.joint_set_param(j, capi.PARAM_FMAX, 0.5)  # steering motor on axis 1
            for w in range(W):
                be.joint_set_param(j, capi.PARAM_VEL, 0.2 * (w - 3) * (1 if jn < 2 else 0), w)
                be.joint_set_param(j, capi.PARAM_VEL | capi.PARAM_GROUP2, 1.0 + 0.5 * w, w)
            if jn == 3:
                be.joint_set_param(j, capi.PARAM_LO_STOP, -0.1)
                be.joint_set_param(j, capi.PARAM_HI_STOP, 0.1)
    for s in range(300):
        for be in (ar, oc):
            be.contacts_clear()
            for w in range(W):
                scenes.sphere_plane_contacts(be, [1, 2, 3, 4], 0.1, w, mu=1.0)
        ar.step(H)
        oc.step(H)
    _compare(ar, oc, 5, tol_pos=5e-4, tol_vel=5e-2)
    for w in (0, 4, 7):
        for j in range(4):
            assert abs(ar.joint_get_angle1(j, w) - oc.joint_get_angle1(j, w)) < 5e-4
            assert abs(ar.joint_get_angle2_rate(j, w) - oc.joint_get_angle2_rate(j, w)) < 5e-2
            assert ar.joint_get_num_rows(j, w) == oc.joint_get_num_rows(j, w)
    assert abs(oc.joint_get_angle2(0, 3)) > 0.05  # the wheels turned (measureAngle2)
    assert abs(ar.joint_get_angle2(0, 3) - oc.joint_get_angle2(0, 3)) < 2e-3


def test_external_forces_and_add_torque(cuda_lib):
    W = 6
    ar, oc = _pair(W)
    for be in (ar, oc):
        sc = scenes.pendulum_chain(be, 2)
    for s in range(150):
        for be in (ar, oc):
            for w in range(W):
                be.body_add_force(1, (0.1 * w, 0.0, 0.5), w)
                be.body_add_force_at_pos(0, (0.0, 0.2, 0.0), (0.1, 0.0, 1.0), w)
                be.joint_add_torque(1, 0.05 * (w - 2), w)
            be.body_add_torque(1, (0.0, 0.01, 0.0))
        ar.step(H)
        oc.step(H)
    _compare(ar, oc, 2)
    assert np.allclose(ar.body_get_force(1, 2), 0.0)  # accumulators are zeroed by the step


def test_ball_and_universal_parity(cuda_lib):
    """Ball (3 rows) and universal (4 rows + motor on axis 1 + stops on axis 2) joints: row counts and RNG state
    exact, poses and the universal's two angles / rates against the fp64 oracle over 300 steps, per-world
    initial spin so that worlds reach the stops at different times."""
    W = 40
    ar, oc = Arena(W), orc.OracleBatch(W)
    scs = [scenes.ball_universal_arm(be) for be in (ar, oc)]
    sc = scs[0]
    for w in range(W):
        for be in (ar, oc):
            be.body_set_angular_vel(sc["bodies"][1], (0.0, 0.2 * (w % 5), 3.0 - 0.15 * w), w)
    ju = sc["universal"]
    hit_stop = 0
    for step in range(300):
        ar.step(H)
        oc.step(H)
        if step % 50 == 49:
            for w in range(0, W, 3):
                assert ar.last_num_rows(w) == oc.last_num_rows(w), (step, w)
                assert ar.joint_get_num_rows(ju, w) == oc.joint_get_num_rows(ju, w), (step, w)
                assert ar.rand_get_seed(w) == oc.rand_get_seed(w)
                hit_stop += oc.joint_get_num_rows(ju, w) == 6
                assert abs(ar.joint_get_angle1(ju, w) - oc.joint_get_angle1(ju, w)) < 2e-4
                assert abs(ar.joint_get_angle2(ju, w) - oc.joint_get_angle2(ju, w)) < 2e-4
                assert abs(ar.joint_get_angle1_rate(ju, w) - oc.joint_get_angle1_rate(ju, w)) < 5e-3
                assert abs(ar.joint_get_angle2_rate(ju, w) - oc.joint_get_angle2_rate(ju, w)) < 5e-3
    assert hit_stop > 0  # some worlds ran into the second axis' stops (motor row + limit row)
    sa = ar.state_download(2).astype(np.float64)
    so = oc.state_download(2)
    d = np.abs(sa - so)
    assert d[:, 0:3].max() < 2e-4 and d[:, 3:7].max() < 5e-4, (d[:, 0:3].max(), d[:, 3:7].max())
    assert oc.joint_get_num_rows(sc["ball"], 0) == 3 and ar.joint_get_num_rows(sc["ball"], 0) == 3
    # the motor drives the first angle at its commanded rate in most worlds
    rates = [oc.joint_get_angle1_rate(ju, w) for w in range(W)]
    assert np.median(np.abs(np.array(rates) - 0.8)) < 0.1, rates
    # device-side observation array agrees with the host-side getter
    fb = ar.joint_feedback_download(2)
    assert np.allclose(fb[ju, :, 5], ar.joint_get_feedback(ju, 5), rtol=0, atol=0)
    assert np.allclose(fb[ju, :, 5], oc.joint_get_feedback(ju, 5), rtol=5e-3, atol=5e-3)
    obs = ar.joint_obs_download(2)
    assert abs(obs[ju, 0, 3] - ar.joint_get_angle1(ju, 3)) < 1e-5 and abs(obs[ju, 2, 3] - ar.joint_get_angle2(ju, 3)) < 1e-5


def test_arm_against_committed_golden(cuda_lib):
    """Device vs the committed oracle fixture (tests/golden/arm_ball_universal_oracle_f64.npz): row counts per
    step and RNG state exact, state and universal angles within fp32 tolerance."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "arm_ball_universal_oracle_f64.npz"))
    W = 6
    ar = Arena(W)
    sc = scenes.ball_universal_arm(ar)
    for w in range(W):
        ar.body_set_angular_vel(sc["bodies"][1], (0.0, 0.3 * w, 3.0 - 0.5 * w), w)
    for k in range(300):
        ar.step(H)
        if k % 25 == 0 or k == 299:
            assert [ar.last_num_rows(w) for w in range(W)] == list(g["rows"][k]), k
    assert [ar.rand_get_seed(w) for w in range(W)] == list(g["seeds"])
    d = np.abs(ar.state_download(2).astype(np.float64) - g["state"])
    assert d[:, 0:3].max() < 2e-4 and d[:, 3:7].max() < 5e-4, (d[:, 0:3].max(), d[:, 3:7].max())
    ju = sc["universal"]
    ang = np.array([[ar.joint_get_angle1(ju, w), ar.joint_get_angle1_rate(ju, w), ar.joint_get_angle2(ju, w),
                     ar.joint_get_angle2_rate(ju, w)] for w in range(W)])
    assert np.abs(ang[:, [0, 2]] - g["angles"][:, [0, 2]]).max() < 5e-4
    assert np.abs(ang[:, [1, 3]] - g["angles"][:, [1, 3]]).max() < 2e-2
