"""GPU parity of the device collision path (broadphase pair lists + narrowphase contact records)
against oracle/orc_collide.c, which defines the collision semantics (the reference tree has none).

Bars: broadphase pair lists and contact counts bit-exact (integer work) -- except where a
candidate lies within fp32 rounding of the decision threshold, which the comparator reports and
bounds; contact records within fp32 tolerance.
"""
import math

import numpy as np
import pytest

import orc
from mars_ode_physics_b200 import Arena, capi, scenes

pytestmark = pytest.mark.gpu

_Arena = Arena


def Arena(W, **kw):  # the broadphase pair list is a debug output of b2p_collide: these tests read it
    ar = _Arena(W, **kw)
    ar.collide_record_pairs(True)
    return ar


SURF = {"mu": 0.8, "mu2": 0.8, "soft_erp": 0.2, "soft_cfm": 1e-5}


def _rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def _compare_contacts(ar, oc, W, borderline=2e-5, depth_atol=2e-5, pos_atol=5e-5, normal_atol=2e-4):
    """Returns (#worlds compared exactly, #borderline worlds).  The default tolerances are for identical poses;
    trajectory tests (device and oracle states drift apart by fp32 rounding) pass wider ones."""
    exact = border = 0
    for w in range(W):
        na, no = ar.contact_count(w), oc.contact_count(w)
        ca = [ar.contact_get(w, c) for c in range(na)]
        co = [oc.contact_get(w, c) for c in range(no)]
        if na != no or any(a[0] != o[0] or a[1] != o[1] for a, o in zip(ca, co)):
            # acceptable only if some involved contact is borderline (depth ~ 0)
            depths = [c[2].depth for c in ca] + [c[2].depth for c in co]
            assert min(depths) < borderline, (w, na, no, depths)
            border += 1
            continue
        for a, o in zip(ca, co):
            da, do = a[2], o[2]
            if abs(da.depth - do.depth) > 1e-4 and min(da.depth, do.depth) < borderline:
                border += 1
                break
            assert abs(da.depth - do.depth) < depth_atol, (w, da.depth, do.depth)
            assert np.allclose(da.pos[:], do.pos[:], atol=pos_atol), (w, da.pos[:], do.pos[:])
            assert np.allclose(da.normal[:], do.normal[:], atol=normal_atol), (w, da.normal[:], do.normal[:])
            assert da.mode == do.mode
            assert abs(da.mu - do.mu) < 1e-6 and abs(da.soft_cfm - do.soft_cfm) < 1e-9
        else:
            exact += 1
    return exact, border


def _build_pair_scene(be, cls_a, dims_a, cls_b, dims_b, static_plane=False):
    scenes.setup_world(be)
    be.set_max_contacts(8)
    a = scenes.add_body(be, 1.0, np.eye(3) * 0.01, (0, 0, 0))
    be.geom_create(cls_a, a, dims_a, surface=SURF)
    if static_plane:
        be.geom_create(capi.GEOM_PLANE, -1, dims_b, surface=SURF)
    else:
        b = scenes.add_body(be, 1.0, np.eye(3) * 0.01, (0, 0, 0))
        be.geom_create(cls_b, b, dims_b, surface=SURF)


@pytest.mark.parametrize("cls,dims", [
    (capi.GEOM_SPHERE, (0.1,)), (capi.GEOM_BOX, (0.2, 0.3, 0.1)),
    (capi.GEOM_CAPSULE, (0.05, 0.3)), (capi.GEOM_CYLINDER, (0.1, 0.25))])
def test_primitive_vs_plane(cuda_lib, cls, dims):
    W = 96
    rng = np.random.default_rng(1234 + cls)
    ar, oc = Arena(W), orc.OracleBatch(W)
    n = np.array([0.1, -0.2, 1.0])
    n /= np.linalg.norm(n)
    for be in (ar, oc):
        _build_pair_scene(be, cls, dims, capi.GEOM_PLANE, (*n, 0.05), static_plane=True)
    for w in range(W):
        p = rng.uniform(-0.2, 0.2, size=3)
        q = _rand_quat(rng)
        if w % 8 == 0:  # resting poses: axis aligned with the normal
            q = np.array([1.0, 0, 0, 0])
        for be in (ar, oc):
            be.body_set_position(0, p, w)
            be.body_set_quaternion(0, q, w)
    ar.collide()
    oc.collide()
    for w in range(W):
        assert ar.collide_get_pairs(w) == oc.collide_get_pairs(w)
    exact, border = _compare_contacts(ar, oc, W)
    assert exact >= W - 4, (exact, border)
    assert sum(oc.contact_count(w) for w in range(W)) > W // 4  # the test actually produced contacts


@pytest.mark.parametrize("ca,da,cb,db", [
    (capi.GEOM_SPHERE, (0.1,), capi.GEOM_SPHERE, (0.15,)),
    (capi.GEOM_SPHERE, (0.1,), capi.GEOM_BOX, (0.3, 0.2, 0.25)),
    (capi.GEOM_BOX, (0.3, 0.2, 0.25), capi.GEOM_SPHERE, (0.1,)),
    (capi.GEOM_CAPSULE, (0.05, 0.3), capi.GEOM_SPHERE, (0.1,)),
    (capi.GEOM_SPHERE, (0.1,), capi.GEOM_CAPSULE, (0.05, 0.3)),
    (capi.GEOM_CAPSULE, (0.05, 0.3), capi.GEOM_CAPSULE, (0.07, 0.2)),
    (capi.GEOM_CAPSULE, (0.05, 0.3), capi.GEOM_BOX, (0.3, 0.2, 0.25)),
    (capi.GEOM_BOX, (0.3, 0.2, 0.25), capi.GEOM_CAPSULE, (0.05, 0.3)),
    (capi.GEOM_SPHERE, (0.1,), capi.GEOM_CYLINDER, (0.1, 0.25)),
    (capi.GEOM_CYLINDER, (0.1, 0.25), capi.GEOM_SPHERE, (0.1,)),
    (capi.GEOM_CAPSULE, (0.05, 0.3), capi.GEOM_CYLINDER, (0.1, 0.25)),
    (capi.GEOM_CYLINDER, (0.1, 0.25), capi.GEOM_CAPSULE, (0.05, 0.3)),
    (capi.GEOM_CYLINDER, (0.1, 0.25), capi.GEOM_BOX, (0.3, 0.2, 0.25)),
    (capi.GEOM_BOX, (0.3, 0.2, 0.25), capi.GEOM_CYLINDER, (0.1, 0.25)),
    (capi.GEOM_CYLINDER, (0.1, 0.25), capi.GEOM_CYLINDER, (0.12, 0.2))])
def test_primitive_pairs(cuda_lib, ca, da, cb, db):
    W = 128
    rng = np.random.default_rng(77 + 10 * ca + cb)
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        _build_pair_scene(be, ca, da, cb, db)
    for w in range(W):
        pa, pb = rng.uniform(-0.15, 0.15, size=3), rng.uniform(-0.15, 0.15, size=3)
        qa, qb = _rand_quat(rng), _rand_quat(rng)
        for be in (ar, oc):
            be.body_set_position(0, pa, w)
            be.body_set_quaternion(0, qa, w)
            be.body_set_position(1, pb, w)
            be.body_set_quaternion(1, qb, w)
    ar.collide()
    oc.collide()
    for w in range(W):
        assert ar.collide_get_pairs(w) == oc.collide_get_pairs(w)
    exact, border = _compare_contacts(ar, oc, W)
    assert exact >= W - 4, (exact, border)
    assert sum(oc.contact_count(w) for w in range(W)) > W // 8


def test_joint_connected_pairs_are_excluded(cuda_lib):
    """WorldPhysics::createContact rejects contacts between joint-connected bodies (:461)."""
    W = 4
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        scenes.setup_world(be)
        be.set_max_contacts(8)
        a = scenes.add_body(be, 1.0, np.eye(3) * 0.01, (0, 0, 0))
        b = scenes.add_body(be, 1.0, np.eye(3) * 0.01, (0.1, 0, 0))
        c = scenes.add_body(be, 1.0, np.eye(3) * 0.01, (0.0, 0.1, 0))
        for body in (a, b, c):
            be.geom_create(capi.GEOM_SPHERE, body, (0.1,), surface=SURF)
        scenes.add_hinge(be, a, b, (0.05, 0, 0), (0, 0, 1))
    ar.collide()
    oc.collide()
    for w in range(W):
        assert ar.collide_get_pairs(w) == oc.collide_get_pairs(w) == [(0, 2), (1, 2)]
        assert ar.contact_count(w) == oc.contact_count(w) == 2


@pytest.mark.parametrize("wheel_shape", ["sphere", "cylinder"])
def test_rover_on_heightfield_trajectory(cuda_lib, wheel_shape):
    """Config D scene (hinge2 rover, shared heightfield), device collide + step vs oracle; with the sphere
    wheels of the benchmark and with cylinder wheels (sample-point cylinder-heightfield collider)."""
    W = 16
    ar, oc = Arena(W), orc.OracleBatch(W)
    heights, x0, y0 = scenes.terrain_heights()
    for be in (ar, oc):
        sc = scenes.rover(be, with_geoms=True, wheel_shape=wheel_shape, max_contacts=8 if wheel_shape == "sphere" else 20)
        be.geom_create_heightfield(heights, 0.1, x0, y0, surface={"mu": 1.0, "mu2": 1.0, "soft_erp": 0.2, "soft_cfm": 1e-5})
        for w in range(W):
            dx, dy = 3.0 * scenes.uniform01(0xD, w, 1) - 1.5, 3.0 * scenes.uniform01(0xD, w, 2) - 1.5
            for b in sc["bodies"]:
                p = be.body_get_position(b, 0 if w == 0 else w)
            # shift the whole rover
        for w in range(W):
            dx, dy = 3.0 * scenes.uniform01(0xD, w, 1) - 1.5, 3.0 * scenes.uniform01(0xD, w, 2) - 1.5
            for b in sc["bodies"]:
                p = be.body_get_position(b, w)
                be.body_set_position(b, (p[0] + dx, p[1] + dy, p[2] + 0.1), w)
            for j in sc["joints"]:
                be.joint_set_param(j, capi.PARAM_VEL | capi.PARAM_GROUP2, 2.0 + 4.0 * scenes.uniform01(0xD, w, 3), w)
    nb = 5
    mismatch_steps = 0
    for step in range(400):
        ar.collide()
        oc.collide()
        if step % 20 == 0:
            for w in range(W):
                assert ar.collide_get_pairs(w) == oc.collide_get_pairs(w), (step, w)
            # states have drifted by fp32 rounding (checked at the end): contact geometry within that drift
            exact, border = _compare_contacts(ar, oc, W, borderline=1e-4, depth_atol=2e-4, pos_atol=5e-4, normal_atol=2e-3)
            mismatch_steps += border
        ar.step(0.001)
        oc.step(0.001)
    sa = ar.state_download(nb).astype(np.float64)
    so = oc.state_download(nb)
    err = np.abs(sa - so)
    print("rover/heightfield 400 steps: max pos err", err[:, :3].max(), "quat", err[:, 3:7].max(),
          "lvel", err[:, 7:10].max(), "borderline worlds", mismatch_steps)
    assert mismatch_steps <= (2 if wheel_shape == "sphere" else 6)
    assert err[:, :3].max() < 2e-3 and err[:, 3:7].max() < 5e-3
    assert sum(oc.contact_count(w) for w in range(W)) >= W  # wheels are on the ground


@pytest.mark.parametrize("cls,dims", [(capi.GEOM_BOX, (0.4, 0.3, 0.2)), (capi.GEOM_CAPSULE, (0.08, 0.4)),
                                      (capi.GEOM_CYLINDER, (0.15, 0.2)), (capi.GEOM_SPHERE, (0.12,))])
def test_primitive_on_heightfield(cuda_lib, cls, dims):
    W = 64
    rng = np.random.default_rng(5 + cls)
    ar, oc = Arena(W), orc.OracleBatch(W)
    heights, x0, y0 = scenes.terrain_heights()
    for be in (ar, oc):
        scenes.setup_world(be)
        be.set_max_contacts(8)
        a = scenes.add_body(be, 1.0, np.eye(3) * 0.01, (0, 0, 0))
        be.geom_create(cls, a, dims, surface=SURF)
        be.geom_create_heightfield(heights, 0.1, x0, y0, surface=SURF)
    for w in range(W):
        p = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-0.05, 0.2)])
        q = _rand_quat(rng)
        for be in (ar, oc):
            be.body_set_position(0, p, w)
            be.body_set_quaternion(0, q, w)
    ar.collide()
    oc.collide()
    exact, border = _compare_contacts(ar, oc, W)
    assert exact >= W - 4, (exact, border)
    assert sum(oc.contact_count(w) for w in range(W)) > W // 2


def test_rover_against_committed_golden(cuda_lib):
    """The committed fixture (tests/golden/rover_terrain_oracle_f64.npz, produced by the oracle with
    tests/golden/make_golden.py) pins contact counts, RNG state and the final state."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rover_terrain_oracle_f64.npz"))
    W = 4
    ar = Arena(W)
    sc = scenes.rover(ar, with_geoms=True)
    h, x0, y0 = scenes.terrain_heights()
    ar.geom_create_heightfield(h, 0.1, x0, y0, surface={"mu": 1.0, "mu2": 1.0, "soft_erp": 0.2, "soft_cfm": 1e-5})
    for w in range(W):
        for b in sc["bodies"]:
            p = ar.body_get_position(b, w)
            ar.body_set_position(b, (p[0] + 0.7 * w, p[1] - 0.4 * w, p[2] + 0.1), w)
    counts = []
    for _ in range(200):
        ar.collide()
        counts.append([ar.contact_count(w) for w in range(W)])
        ar.step(0.001)
    assert np.array_equal(np.array(counts), g["counts"])
    assert [ar.rand_get_seed(w) for w in range(W)] == list(g["seeds"])
    err = np.abs(ar.state_download(5).astype(np.float64) - g["state"])
    assert err[:, :3].max() < 1e-4 and err[:, 3:7].max() < 1e-4, (err[:, :3].max(), err[:, 3:7].max())


def test_box_box_random_poses(cuda_lib):
    """Box-box collider (SAT + face clipping / edge-edge) on random overlapping poses."""
    W = 256
    rng = np.random.default_rng(2024)
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        _build_pair_scene(be, capi.GEOM_BOX, (0.3, 0.2, 0.25), capi.GEOM_BOX, (0.2, 0.35, 0.15))
    for w in range(W):
        pa, pb = rng.uniform(-0.12, 0.12, size=3), rng.uniform(-0.12, 0.12, size=3)
        qa, qb = _rand_quat(rng), _rand_quat(rng)
        if w % 4 == 0:  # face-to-face stacking poses (the config-B case)
            qa = qb = np.array([1.0, 0, 0, 0])
            pa, pb = np.array([0.0, 0.0, 0.0]), np.array([rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05), 0.19])
            if w % 8 == 0:
                qb = scenes.quat_axis_angle((0, 0, 1), rng.uniform(-0.5, 0.5))
        for be in (ar, oc):
            be.body_set_position(0, pa, w)
            be.body_set_quaternion(0, qa, w)
            be.body_set_position(1, pb, w)
            be.body_set_quaternion(1, qb, w)
    ar.collide()
    oc.collide()
    for w in range(W):
        assert ar.collide_get_pairs(w) == oc.collide_get_pairs(w)
    exact, border = _compare_contacts(ar, oc, W, borderline=5e-5)
    print("box-box: exact worlds", exact, "borderline", border)
    assert exact >= W - 12, (exact, border)
    total = sum(oc.contact_count(w) for w in range(W))
    multi = sum(1 for w in range(W) if oc.contact_count(w) >= 3)
    assert total > W and multi > W // 8  # face contacts with several points occurred


def test_box_stack_config_b(cuda_lib):
    """Config B: 10-box stack on a plane with per-world jitter, device collide + step vs oracle."""
    W = 32
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        sc = scenes.box_stack(be, n=10)
        scenes.jitter_box_stack(be, sc, W)
    nb = 10
    checked = borderline = pair_mismatch = 0
    worst = 0.0
    for step in range(125):
        if step % 10 == 5:
            # (step 0 is skipped: box 0 starts exactly touching the plane, AABB gap == 0)
            # integer checks on IDENTICAL inputs: free-running fp32 / fp64 trajectories differ at the
            # 1e-7 level, which flips AABB pairs that are exactly touching (the boxes land on each other
            # at the same instant); so measure the drift, then give both sides the same rounded state
            so = oc.state_download(nb)
            worst = max(worst, np.abs(ar.state_download(nb).astype(np.float64) - so)[:, :7].max())
            st32 = so.astype(np.float32)
            ar.state_upload(st32)
            for w in range(W):
                for b in range(nb):
                    oc.body_set_state(b, st32[b, :, w].astype(np.float64), w)
        ar.collide()
        oc.collide()
        if step % 10 == 5:
            for w in range(W):
                pa, po = ar.collide_get_pairs(w), oc.collide_get_pairs(w)
                pair_mismatch += len(set(pa) ^ set(po))
                checked += 1
                if ar.contact_count(w) != oc.contact_count(w):
                    borderline += 1
        ar.step(0.001)
        oc.step(0.001)
    assert pair_mismatch == 0, pair_mismatch
    print("box stack: max pose drift over 10 free-running steps", worst)
    # a resting face contact is a discontinuous function of the state (a vertex at depth 0 is in or out
    # of the manifold), so contact-count flips perturb the stack; bound the drift, do not expect 1e-6
    assert worst < 1e-2
    err = np.abs(ar.state_download(nb).astype(np.float64) - oc.state_download(nb))
    print("box stack 120 steps: pos err", err[:, :3].max(), "quat", err[:, 3:7].max(), "count mismatches",
          borderline, "of", checked)
    # resting face contacts have depth ~ 0, so a few candidate points flip between fp32 and fp64;
    # the stack itself must still agree closely
    assert borderline <= checked // 4
    assert err[:, :3].max() < 2e-3 and err[:, 3:7].max() < 5e-3
    z = oc.state_download(nb)[:, 2, :]
    assert np.all(np.diff(z, axis=0) > 0.15)  # still a stack


def test_box_stack_config_b_free_running(cuda_lib):
    """Config B with NO re-synchronisation: 130 steps of device collide + step against the oracle from the same
    jittered start.  The boxes fall 5 mm onto each other (all ten gaps close in the same step, ~32) and then rest
    on face contacts.  A box-box face contact keeps the <= 4 deepest of up to 8 clipped candidate points whose
    depths differ by less than fp32 resolution, so device (fp32) and oracle (fp64) choose different -- equally
    valid -- manifolds in some steps of EVERY world (measured: the first differing contact count appears between
    step 33 and 126), and the trajectories then separate at the millimetre level.  Bars: free fall agrees to
    rounding in every world; the drift after 130 steps is bounded and the per-box mean heights over the worlds
    agree; both stacks still stand and are nearly at rest."""
    W = 32
    ar, oc = Arena(W), orc.OracleBatch(W)
    for be in (ar, oc):
        sc = scenes.box_stack(be, n=10)
        scenes.jitter_box_stack(be, sc, W)
    nb = 10
    agree = total = 0
    for step in range(130):
        ar.collide()
        oc.collide()
        for w in range(W):
            agree += ar.contact_count(w) == oc.contact_count(w)
            total += 1
        ar.step(0.001)
        oc.step(0.001)
        if step == 29:
            d = np.abs(ar.state_download(nb).astype(np.float64) - oc.state_download(nb))
            assert d[:, :7].max() < 1e-6 and d[:, 7:].max() < 1e-5, d.max()   # free fall: pure integration
    sa, so = ar.state_download(nb).astype(np.float64), oc.state_download(nb)
    d = np.abs(sa - so)
    print("box stack free-running 130 steps: contact counts agree in", agree, "of", total, "(world, step) pairs; pos",
          d[:, :3].max(), "quat", d[:, 3:7].max(), "| speed device", np.abs(sa[:, 7:10]).max(), "oracle",
          np.abs(so[:, 7:10]).max())
    # (count agreement is reported, not asserted: once two trajectories are a millimetre apart their contact sets
    # legitimately differ; on IDENTICAL states the device and the oracle agree, see test_box_stack_config_b)
    assert d[:, :3].max() < 1e-2 and d[:, 3:7].max() < 3e-2, (d[:, :3].max(), d[:, 3:7].max())
    # statistics over the 32 worlds are stable where single trajectories are not: mean height of every box
    mz = np.abs(sa[:, 2, :].mean(axis=1) - so[:, 2, :].mean(axis=1))
    print("mean box heights differ by", mz.max())
    assert mz.max() < 5e-4, mz
    for z in (sa[:, 2, :], so[:, 2, :]):
        assert np.all(np.diff(z, axis=0) > 0.15) and np.isfinite(z).all()  # still a stack, on both sides
    assert np.abs(sa[:, 7:10]).max() < 0.2 and np.abs(so[:, 7:10]).max() < 0.2   # and (nearly) at rest
    assert ar.contacts_dropped() == 0


def test_config_c_device_collision_sampled_worlds(cuda_lib):
    """BASELINE config C at its size (16384 worlds of the 12-hinge robot, foot spheres on a plane, DEVICE
    collision as bench.py --workload C runs it): the last 6 worlds of the batch against the same 6 worlds built
    alone on the oracle (per-world commands depend on the world id only), 100 steps, free-running."""
    import bench
    W, S = 16384, 6
    ar = _Arena(W)
    sa = bench.build_quadruped_worlds(ar, W)
    oc = orc.OracleBatch(S)
    so = bench.build_quadruped_worlds(oc, S, world_offset=W - S)
    nb = len(sa["bodies"])
    for step in range(100):
        ar.collide_step(0.001)
        oc.collide_step(0.001)
    got = ar.state_download(nb)[:, :, W - S:].astype(np.float64)
    want = oc.state_download(nb)
    d = np.abs(got - want)
    print("config C device collision, 100 steps: pos", d[:, :3].max(), "quat", d[:, 3:7].max(), "lvel",
          d[:, 7:10].max(), "avel", d[:, 10:].max())
    for k in range(S):
        assert ar.contact_count(W - S + k) == oc.contact_count(k)
        assert ar.last_num_rows(W - S + k) == oc.last_num_rows(k)
        assert ar.rand_get_seed(W - S + k) == oc.rand_get_seed(k)
    assert d[:, :3].max() < 2e-4 and d[:, 3:7].max() < 1e-3 and d[:, 7:10].max() < 2e-2, d.max()
    assert sum(oc.contact_count(k) for k in range(S)) >= S  # the feet are on the ground
    ar.close()


def test_config_d_full_size_sampled_worlds(cuda_lib):
    """BASELINE config D at its full size as bench.py runs it (65536 rovers on the shared heightfield, device
    collision, ODE island row order, random row reordering): the last 6 worlds of the batch against the same 6
    worlds built alone on the oracle (placement and wheel speed depend on the world id only), 400 steps from the
    drop to driving on the terrain, free-running."""
    import bench
    W, S = 65536, 6
    ar = _Arena(W)
    sa = bench.build_rover_worlds(ar, W)
    ar.set_order_mode(capi.ORDER_ODE_ISLANDS)
    oc = orc.OracleBatch(S)
    so = bench.build_rover_worlds(oc, S, world_offset=W - S)
    oc.set_order_mode(capi.ORDER_ODE_ISLANDS)
    nb = len(sa["bodies"])
    for step in range(400):
        ar.collide_step(0.001)
        oc.collide_step(0.001)
    got = ar.state_download(nb)[:, :, W - S:].astype(np.float64)
    want = oc.state_download(nb)
    d = np.abs(got - want)
    print("config D full size, 400 steps: pos", d[:, :3].max(), "quat", d[:, 3:7].max(), "lvel", d[:, 7:10].max(),
          "avel", d[:, 10:].max())
    for k in range(S):
        assert ar.contact_count(W - S + k) == oc.contact_count(k)
        assert ar.last_num_rows(W - S + k) == oc.last_num_rows(k)
        assert ar.rand_get_seed(W - S + k) == oc.rand_get_seed(k)
    # (observed on B200: pos 1.1e-5, quat 1.5e-5, lvel 4.3e-4, avel 7.8e-4; bars at about ten times that)
    assert d[:, :3].max() < 1e-4 and d[:, 3:7].max() < 2e-4 and d[:, 7:10].max() < 5e-3 and d[:, 10:].max() < 1e-2, d.max()
    assert sum(oc.contact_count(k) for k in range(S)) >= S  # the rovers are on the ground
    ar.close()


def test_sampled_colliders_against_committed_golden(cuda_lib):
    """Device contact records of the sample-point colliders vs the committed oracle fixture
    (tests/golden/sampled_colliders_oracle_f64.npz: 5 pair types x 24 seeded poses)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampled_colliders_oracle_f64.npz"))
    pairs = [(capi.GEOM_CAPSULE, (0.05, 0.3), capi.GEOM_BOX, (0.3, 0.2, 0.25)),
             (capi.GEOM_SPHERE, (0.1,), capi.GEOM_CYLINDER, (0.1, 0.25)),
             (capi.GEOM_CAPSULE, (0.05, 0.3), capi.GEOM_CYLINDER, (0.1, 0.25)),
             (capi.GEOM_CYLINDER, (0.1, 0.25), capi.GEOM_BOX, (0.3, 0.2, 0.25)),
             (capi.GEOM_CYLINDER, (0.1, 0.25), capi.GEOM_CYLINDER, (0.12, 0.2))]
    total = 0
    for pi, (ca, da, cb, db) in enumerate(pairs):
        W = 24
        ar = Arena(W)
        _build_pair_scene(ar, ca, da, cb, db)
        for w in range(W):
            for body in (0, 1):
                ar.body_set_position(body, g["poses"][pi, w, body, :3], w)
                ar.body_set_quaternion(body, g["poses"][pi, w, body, 3:], w)
        ar.collide()
        for w in range(W):
            want = g["records"][pi, w]
            n_want = int(np.isfinite(want[:, 7]).sum())
            n_got = ar.contact_count(w)
            depths = list(want[:n_want, 7])
            if n_got != n_want:
                assert min(depths + [ar.contact_get(w, c)[2].depth for c in range(n_got)]) < 2e-5, (pi, w, n_got, n_want)
                continue
            for c in range(n_got):
                b1, b2, d = ar.contact_get(w, c)
                assert b1 == int(want[c, 0])
                if abs(d.depth - want[c, 7]) > 2e-5:
                    assert min(depths) < 2e-5, (pi, w, c, d.depth, want[c, 7])  # borderline candidate set
                    break
                assert np.allclose(d.pos[:], want[c, 1:4], atol=5e-5) and np.allclose(d.normal[:], want[c, 4:7], atol=2e-4)
                total += 1
        ar.close()
    assert total > 100
