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


def _run(path, t_end, worlds=1):
    assert os.path.exists(BIN), "facade test binary missing: run __graft_entry__.build()"
    subprocess.run([BIN, path, str(t_end), str(worlds)], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
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
    rows = orc_lib.orc_mars_run_reproducable_test(ocsv.encode(), 0.5, orc.ORDER_WORLD_BLOCK, None)
    ora = np.loadtxt(ocsv, delimiter=",", skiprows=1)
    assert rows == len(ora) == len(gpu)
    assert np.array_equal(gpu[:, 0], ora[:, 0])  # identical time column (same double accumulation)
    err = np.abs(gpu[:, 1:] - ora[:, 1:])
    print("facade vs oracle: max |dx| over 0.1 s", err[:100].max(), " over 0.5 s", err.max())
    assert err[:100].max() < 2e-4
    assert err.max() < 5e-3
    assert np.abs(ora[-1, 1:]).max() > 1e-3  # the robot actually moved


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
