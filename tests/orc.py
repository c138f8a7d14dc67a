"""ctypes access to the CPU oracle (oracle/build/liborc_f64.so / liborc_f32.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import this module.  It is never on the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

JOINT_HINGE, JOINT_HINGE2, JOINT_SLIDER, JOINT_FIXED, JOINT_CONTACT = 1, 2, 3, 6, 100
ORDER_WORLD_BLOCK, ORDER_ODE_ISLANDS = 0, 1
REORDER_RANDOMLY, REORDER_NONE = 0, 1


class ContactDesc(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("normal", C.c_double * 3), ("depth", C.c_double),
                ("mode", C.c_int), ("mu", C.c_double), ("mu2", C.c_double), ("rho", C.c_double),
                ("rho2", C.c_double), ("rhoN", C.c_double), ("bounce", C.c_double),
                ("bounce_vel", C.c_double), ("soft_erp", C.c_double), ("soft_cfm", C.c_double),
                ("motion1", C.c_double), ("motion2", C.c_double), ("motionN", C.c_double),
                ("slip1", C.c_double), ("slip2", C.c_double)]


class Surface(C.Structure):
    _fields_ = [("mu", C.c_double), ("mu2", C.c_double), ("rho", C.c_double), ("rho2", C.c_double),
                ("rhoN", C.c_double), ("soft_erp", C.c_double), ("soft_cfm", C.c_double)]


class Mass(C.Structure):
    _fields_ = [("mass", C.c_double), ("c", C.c_double * 3), ("I", C.c_double * 9)]


class MarsContact(C.Structure):
    _fields_ = [("frame1", C.c_int), ("frame2", C.c_int), ("pos", C.c_double * 3),
                ("normal", C.c_double * 3), ("depth", C.c_double), ("cfm", C.c_double),
                ("erp", C.c_double), ("friction1", C.c_double), ("friction2", C.c_double),
                ("rolling_friction", C.c_double), ("rolling_friction2", C.c_double),
                ("spinning_friction", C.c_double), ("bounce", C.c_double), ("bounce_vel", C.c_double),
                ("motion1", C.c_double), ("motion2", C.c_double), ("fds1", C.c_double),
                ("fds2", C.c_double)]


class MarsJointCfg(C.Structure):
    _fields_ = [("type", C.c_int), ("parent", C.c_int), ("child", C.c_int),
                ("anchor", C.c_double * 3), ("axis1", C.c_double * 3), ("axis2", C.c_double * 3),
                ("has_spring_damping", C.c_int), ("has_spring_stiffness", C.c_int),
                ("has_joint_cfm", C.c_int), ("has_friction", C.c_int),
                ("spring_damping", C.c_double), ("spring_stiffness", C.c_double),
                ("joint_cfm", C.c_double), ("friction_coefficient", C.c_double)]


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_libs = {}


def load(single=False):
    if single in _libs:
        return _libs[single]
    path = os.path.join(ORACLE_DIR, "build", "liborc_f32.so" if single else "liborc_f64.so")
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    P, D, I = C.c_void_p, C.c_double, C.c_int
    DP = C.POINTER(C.c_double)
    sig = {
        "orc_real_size": (I, []),
        "orc_world_create": (P, []), "orc_world_destroy": (None, [P]),
        "orc_world_set_gravity": (None, [P, D, D, D]), "orc_world_set_erp": (None, [P, D]),
        "orc_world_set_cfm": (None, [P, D]), "orc_world_set_max_angular_speed": (None, [P, D]),
        "orc_world_set_contact_max_correcting_vel": (None, [P, D]),
        "orc_world_set_contact_surface_layer": (None, [P, D]),
        "orc_world_set_quickstep_num_iterations": (None, [P, I]),
        "orc_world_set_quickstep_w": (None, [P, D]), "orc_world_set_order_mode": (None, [P, I]),
        "orc_world_set_reorder_method": (None, [P, I]),
        "orc_world_rand_set_seed": (None, [P, C.c_uint32]), "orc_world_rand_get_seed": (C.c_uint32, [P]),
        "orc_world_set_variant": (None, [P, I, I]),
        "orc_world_set_stepper": (None, [P, I]), "orc_world_last_lcp_error": (I, [P]),
        "orc_world_quickstep": (None, [P, D]), "orc_world_last_num_rows": (I, [P]),
        "orc_world_last_num_islands": (I, [P]), "orc_world_last_residual": (D, [P]),
        "orc_world_num_bodies": (I, [P]), "orc_world_num_joints": (I, [P]),
        "orc_world_num_contacts": (I, [P]),
        "orc_rand_next": (C.c_uint32, [C.POINTER(C.c_uint32)]),
        "orc_rand_int": (I, [C.POINTER(C.c_uint32), I]),
        "orc_body_create": (I, [P]), "orc_body_set_mass": (None, [P, I, D, DP]),
        "orc_body_get_mass": (None, [P, I, DP, DP]),
        "orc_body_set_position": (None, [P, I, D, D, D]),
        "orc_body_set_quaternion": (None, [P, I, D, D, D, D]),
        "orc_body_set_linear_vel": (None, [P, I, D, D, D]),
        "orc_body_set_angular_vel": (None, [P, I, D, D, D]),
        "orc_body_set_force": (None, [P, I, D, D, D]), "orc_body_set_torque": (None, [P, I, D, D, D]),
        "orc_body_add_force": (None, [P, I, D, D, D]), "orc_body_add_torque": (None, [P, I, D, D, D]),
        "orc_body_add_force_at_pos": (None, [P, I, D, D, D, D, D, D]),
        "orc_body_set_gravity_mode": (None, [P, I, I]), "orc_body_set_gyroscopic_mode": (None, [P, I, I]),
        "orc_body_get_position": (None, [P, I, DP]), "orc_body_get_quaternion": (None, [P, I, DP]),
        "orc_body_get_rotation": (None, [P, I, DP]), "orc_body_get_linear_vel": (None, [P, I, DP]),
        "orc_body_get_angular_vel": (None, [P, I, DP]), "orc_body_get_force": (None, [P, I, DP]),
        "orc_body_get_torque": (None, [P, I, DP]), "orc_body_get_state": (None, [P, I, DP]),
        "orc_body_set_state": (None, [P, I, DP]),
        "orc_joint_create": (I, [P, I, I, I]), "orc_joint_set_anchor": (None, [P, I, D, D, D]),
        "orc_joint_set_axis1": (None, [P, I, D, D, D]), "orc_joint_set_axis2": (None, [P, I, D, D, D]),
        "orc_joint_set_fixed": (None, [P, I]), "orc_joint_set_param": (None, [P, I, I, D]),
        "orc_joint_get_param": (D, [P, I, I]), "orc_joint_get_anchor": (None, [P, I, DP]),
        "orc_joint_get_anchor2": (None, [P, I, DP]), "orc_joint_get_axis1": (None, [P, I, DP]),
        "orc_joint_get_axis2": (None, [P, I, DP]), "orc_joint_get_angle1": (D, [P, I]),
        "orc_joint_get_angle1_rate": (D, [P, I]), "orc_joint_get_angle2": (D, [P, I]),
        "orc_joint_get_angle2_rate": (D, [P, I]), "orc_joint_add_torque": (None, [P, I, D]),
        "orc_joint_get_feedback": (None, [P, I, DP]), "orc_joint_get_num_rows": (I, [P, I]),
        "orc_are_connected_excluding": (I, [P, I, I, I]),
        "orc_contacts_clear": (None, [P]), "orc_contact_add": (I, [P, I, I, C.POINTER(ContactDesc)]),
        "orc_contact_get_feedback": (None, [P, I, DP]),
        "orc_geom_create": (I, [P, I, I, DP, DP, DP, C.POINTER(Surface)]),
        "orc_geom_create_heightfield": (I, [P, P, I, I, D, D, D, C.POINTER(Surface)]),
        "orc_world_set_max_contacts": (None, [P, I]),
        "orc_post_joint_update": (None, [P, I, DP]),
        "orc_post_joint_friction": (None, [P, I, D]),
        "orc_post_frame_update": (None, [P, I, D, D, D, DP, DP]),
        "orc_post_body_contact_force": (None, [P, I, DP]),
        "orc_post_step_world": (None, [P, D, P, P, P, P, P, P, P]),
        "orc_collide": (I, [P]),
        "orc_collide_get_pairs": (I, [P, C.POINTER(C.c_int), I]),
        "orc_contact_get": (I, [P, I, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(ContactDesc)]),
        "orc_bench_run": (None, [C.POINTER(P), I, I, D, I]),
        "orc_mass_set_zero": (None, [C.POINTER(Mass)]),
        "orc_mass_set_box_total": (None, [C.POINTER(Mass), D, D, D, D]),
        "orc_mass_set_box": (None, [C.POINTER(Mass), D, D, D, D]),
        "orc_mass_set_sphere_total": (None, [C.POINTER(Mass), D, D]),
        "orc_mass_set_sphere": (None, [C.POINTER(Mass), D, D]),
        "orc_mass_set_capsule_total": (None, [C.POINTER(Mass), D, I, D, D]),
        "orc_mass_set_capsule": (None, [C.POINTER(Mass), D, I, D, D]),
        "orc_mass_set_cylinder_total": (None, [C.POINTER(Mass), D, I, D, D]),
        "orc_mass_set_cylinder": (None, [C.POINTER(Mass), D, I, D, D]),
        "orc_mass_adjust": (None, [C.POINTER(Mass), D]),
        "orc_mass_rotate": (None, [C.POINTER(Mass), DP]),
        "orc_mass_translate": (None, [C.POINTER(Mass), D, D, D]),
        "orc_mass_add": (None, [C.POINTER(Mass), C.POINTER(Mass)]),
        "orc_mass_check": (I, [C.POINTER(Mass)]),
        "orc_plane_space": (None, [DP, DP, DP]), "orc_q_to_R": (None, [DP, DP]),
        # mars glue
        "orc_mars_create": (P, []), "orc_mars_destroy": (None, [P]),
        "orc_mars_init_the_world": (None, [P]), "orc_mars_world": (P, [P]),
        "orc_mars_set_step_size": (None, [P, D]), "orc_mars_get_step_size": (D, [P]),
        "orc_mars_clear_previous_step": (None, [P]), "orc_mars_step_the_world": (None, [P]),
        "orc_mars_create_frame": (I, [P, D, D]),
        "orc_mars_frame_add_object": (None, [P, I, C.POINTER(Mass), DP, DP]),
        "orc_mars_frame_body": (I, [P, I]), "orc_mars_frame_acquire_body": (I, [P, I]),
        "orc_mars_frame_get_position": (None, [P, I, DP]),
        "orc_mars_frame_set_position": (None, [P, I, DP]),
        "orc_mars_frame_get_rotation": (None, [P, I, DP]),
        "orc_mars_frame_set_rotation": (None, [P, I, DP]),
        "orc_mars_frame_get_linear_acceleration": (None, [P, I, DP]),
        "orc_mars_frame_get_contact_force": (None, [P, I, DP]),
        "orc_mars_frame_get_mass": (None, [P, I, DP, DP]),
        "orc_mars_frame_add_contact": (I, [P, I, C.POINTER(MarsContact)]),
        "orc_mars_create_joint": (I, [P, C.POINTER(MarsJointCfg)]),
        "orc_mars_joint_set_force_limit": (None, [P, I, D]),
        "orc_mars_joint_set_velocity": (None, [P, I, D]),
        "orc_mars_joint_set_velocity2": (None, [P, I, D]),
        "orc_mars_joint_set_joint_as_motor": (None, [P, I, I]),
        "orc_mars_joint_get_position": (D, [P, I]), "orc_mars_joint_get_velocity": (D, [P, I]),
        "orc_mars_joint_get_motor_torque": (D, [P, I]),
        "orc_mars_joint_get_axis_torque": (None, [P, I, DP]),
        "orc_mars_joint_get_joint_load": (None, [P, I, DP]),
        "orc_mars_joint_ode_id": (I, [P, I]),
        "orc_mars_build_reproducable_scene": (None, [P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "orc_mars_reproducable_tick": (I, [P, C.POINTER(C.c_int), C.POINTER(C.c_int), D]),
        "orc_mars_run_reproducable_test": (I, [C.c_char_p, D, I, DP]),
        "orc_mars_run_reproducable_test_stepper": (I, [C.c_char_p, D, I, I, DP]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _libs[single] = lib
    return lib


def _d(n):
    return (C.c_double * n)()


class OracleBatch:
    """W independent oracle worlds behind the method names of mars_ode_physics_b200.Arena."""

    def __init__(self, num_worlds, single=False):
        self.lib = load(single)
        self.num_worlds = num_worlds
        self.worlds = [self.lib.orc_world_create() for _ in range(num_worlds)]
        self.max_contacts = 8

    def close(self):
        for w in self.worlds:
            self.lib.orc_world_destroy(w)
        self.worlds = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _sel(self, world):
        return self.worlds if world == -1 else [self.worlds[world]]

    def _all(self, fn, *a):
        r = None
        for w in self.worlds:
            r = fn(w, *a)
        return r

    # world
    def set_max_contacts(self, n):
        self.max_contacts = n
        self._all(self.lib.orc_world_set_max_contacts, n)

    def enable_rolling_friction(self, on=True):
        pass

    def set_gravity(self, x, y, z):
        self._all(self.lib.orc_world_set_gravity, x, y, z)

    def set_cfm(self, v):
        self._all(self.lib.orc_world_set_cfm, v)

    def set_erp(self, v):
        self._all(self.lib.orc_world_set_erp, v)

    def set_max_angular_speed(self, v):
        self._all(self.lib.orc_world_set_max_angular_speed, v)

    def set_contact_max_correcting_vel(self, v):
        self._all(self.lib.orc_world_set_contact_max_correcting_vel, v)

    def set_contact_surface_layer(self, v):
        self._all(self.lib.orc_world_set_contact_surface_layer, v)

    def set_quickstep_num_iterations(self, n):
        self._all(self.lib.orc_world_set_quickstep_num_iterations, n)

    def set_quickstep_w(self, v):
        self._all(self.lib.orc_world_set_quickstep_w, v)

    def set_reorder_method(self, m):
        self._all(self.lib.orc_world_set_reorder_method, m)

    def set_order_mode(self, m):
        self._all(self.lib.orc_world_set_order_mode, m)

    def set_stepper(self, stepper):
        """0: dWorldQuickStep (SOR-LCP); 1: dWorldStep (dense LCP, Dantzig)"""
        self._all(self.lib.orc_world_set_stepper, stepper)

    def last_lcp_error(self, world):
        return self.lib.orc_world_last_lcp_error(self.worlds[world])

    def set_variant(self, which, value):
        """restatement choices (orc.h ORC_VARIANT_*)"""
        self._all(self.lib.orc_world_set_variant, which, value)

    def rand_set_seed(self, seed):
        self._all(self.lib.orc_world_rand_set_seed, seed)

    def rand_get_seed(self, world):
        return self.lib.orc_world_rand_get_seed(self.worlds[world])

    # bodies
    def body_create(self):
        return self._all(self.lib.orc_body_create)

    def body_set_mass(self, b, mass, I):
        arr = (C.c_double * 9)(*[float(x) for x in np.asarray(I, dtype=float).reshape(9)])
        self._all(self.lib.orc_body_set_mass, b, mass, arr)

    def _set(self, fn, b, v, world):
        for w in self._sel(world):
            fn(w, b, *map(float, v))

    def body_set_position(self, b, p, world=-1):
        self._set(self.lib.orc_body_set_position, b, p, world)

    def body_set_quaternion(self, b, q, world=-1):
        self._set(self.lib.orc_body_set_quaternion, b, q, world)

    def body_set_linear_vel(self, b, v, world=-1):
        self._set(self.lib.orc_body_set_linear_vel, b, v, world)

    def body_set_angular_vel(self, b, v, world=-1):
        self._set(self.lib.orc_body_set_angular_vel, b, v, world)

    def body_set_force(self, b, v, world=-1):
        self._set(self.lib.orc_body_set_force, b, v, world)

    def body_set_torque(self, b, v, world=-1):
        self._set(self.lib.orc_body_set_torque, b, v, world)

    def body_add_force(self, b, v, world=-1):
        self._set(self.lib.orc_body_add_force, b, v, world)

    def body_add_torque(self, b, v, world=-1):
        self._set(self.lib.orc_body_add_torque, b, v, world)

    def body_add_force_at_pos(self, b, f, p, world=-1):
        for w in self._sel(world):
            self.lib.orc_body_add_force_at_pos(w, b, *map(float, f), *map(float, p))

    def body_set_gravity_mode(self, b, on):
        self._all(self.lib.orc_body_set_gravity_mode, b, int(on))

    def body_set_gyroscopic_mode(self, b, on):
        self._all(self.lib.orc_body_set_gyroscopic_mode, b, int(on))

    def _get(self, fn, n, world, *ids):
        out = _d(n)
        fn(self.worlds[world], *ids, out)
        return np.array(out[:], dtype=float)

    def body_get_position(self, b, world=0):
        return self._get(self.lib.orc_body_get_position, 3, world, b)

    def body_get_quaternion(self, b, world=0):
        return self._get(self.lib.orc_body_get_quaternion, 4, world, b)

    def body_get_rotation(self, b, world=0):
        return self._get(self.lib.orc_body_get_rotation, 9, world, b)

    def body_get_linear_vel(self, b, world=0):
        return self._get(self.lib.orc_body_get_linear_vel, 3, world, b)

    def body_get_angular_vel(self, b, world=0):
        return self._get(self.lib.orc_body_get_angular_vel, 3, world, b)

    def body_get_force(self, b, world=0):
        return self._get(self.lib.orc_body_get_force, 3, world, b)

    def body_get_torque(self, b, world=0):
        return self._get(self.lib.orc_body_get_torque, 3, world, b)

    def body_get_state(self, b, world=0):
        return self._get(self.lib.orc_body_get_state, 13, world, b)

    def body_set_state(self, b, s, world=-1):
        arr = (C.c_double * 13)(*map(float, s))
        for w in self._sel(world):
            self.lib.orc_body_set_state(w, b, arr)

    def state_download(self, num_bodies, out=None):
        if out is None:
            out = np.empty((num_bodies, 13, self.num_worlds), dtype=np.float64)
        buf = _d(13)
        for wi, w in enumerate(self.worlds):
            for b in range(num_bodies):
                self.lib.orc_body_get_state(w, b, buf)
                out[b, :, wi] = buf[:]
        return out

    # joints
    def joint_create(self, jtype, b1, b2):
        return self._all(self.lib.orc_joint_create, jtype, b1, b2)

    def joint_set_anchor(self, j, p):
        self._all(self.lib.orc_joint_set_anchor, j, *map(float, p))

    def joint_set_axis1(self, j, a):
        self._all(self.lib.orc_joint_set_axis1, j, *map(float, a))

    def joint_set_axis2(self, j, a):
        self._all(self.lib.orc_joint_set_axis2, j, *map(float, a))

    def joint_set_fixed(self, j):
        self._all(self.lib.orc_joint_set_fixed, j)

    def joint_set_param(self, j, param, value, world=-1):
        for w in self._sel(world):
            self.lib.orc_joint_set_param(w, j, param, float(value))

    def joint_get_param(self, j, param, world=0):
        return self.lib.orc_joint_get_param(self.worlds[world], j, param)

    def joint_set_param_batch(self, j, param, values):
        for w, v in zip(self.worlds, values):
            self.lib.orc_joint_set_param(w, j, param, float(v))

    def joint_get_anchor(self, j, world=0):
        return self._get(self.lib.orc_joint_get_anchor, 3, world, j)

    def joint_get_anchor2(self, j, world=0):
        return self._get(self.lib.orc_joint_get_anchor2, 3, world, j)

    def joint_get_axis1(self, j, world=0):
        return self._get(self.lib.orc_joint_get_axis1, 3, world, j)

    def joint_get_axis2(self, j, world=0):
        return self._get(self.lib.orc_joint_get_axis2, 3, world, j)

    def joint_get_angle1(self, j, world=0):
        return self.lib.orc_joint_get_angle1(self.worlds[world], j)

    def joint_get_angle1_rate(self, j, world=0):
        return self.lib.orc_joint_get_angle1_rate(self.worlds[world], j)

    def joint_get_angle2(self, j, world=0):
        return self.lib.orc_joint_get_angle2(self.worlds[world], j)

    def joint_get_angle2_rate(self, j, world=0):
        return self.lib.orc_joint_get_angle2_rate(self.worlds[world], j)

    def joint_add_torque(self, j, t, world=-1):
        for w in self._sel(world):
            self.lib.orc_joint_add_torque(w, j, float(t))

    def joint_get_feedback(self, j, world=0):
        return self._get(self.lib.orc_joint_get_feedback, 13, world, j)

    def joint_get_num_rows(self, j, world=0):
        return self.lib.orc_joint_get_num_rows(self.worlds[world], j)

    # contacts
    def contacts_clear(self, world=-1):
        for w in self._sel(world):
            self.lib.orc_contacts_clear(w)

    def contact_add(self, world, b1, b2, **kw):
        d = ContactDesc()
        d.pos[:] = [float(x) for x in kw["pos"]]
        d.normal[:] = [float(x) for x in kw["normal"]]
        d.depth = float(kw["depth"])
        d.mode = int(kw["mode"])
        for k in ("mu", "mu2", "rho", "rho2", "rhoN", "soft_erp", "soft_cfm"):
            setattr(d, k, float(kw.get(k, 0.0)))
        return self.lib.orc_contact_add(self.worlds[world], b1, b2, C.byref(d))

    def contact_count(self, world=0):
        return self.lib.orc_world_num_contacts(self.worlds[world])

    def contact_get_feedback(self, world, c):
        out = _d(12)
        self.lib.orc_contact_get_feedback(self.worlds[world], c, out)
        return np.array(out[:3], dtype=float)

    def contact_get(self, world, c):
        b1, b2, d = C.c_int(), C.c_int(), ContactDesc()
        rc = self.lib.orc_contact_get(self.worlds[world], c, C.byref(b1), C.byref(b2), C.byref(d))
        assert rc == 0
        return b1.value, b2.value, d

    # geoms / collision
    @staticmethod
    def _surface(s):
        if s is None:
            return None
        out = Surface()
        for k in ("mu", "mu2", "rho", "rho2", "rhoN", "soft_erp", "soft_cfm"):
            setattr(out, k, float(s.get(k, 0.0)))
        return out

    def geom_create(self, cls, body, dims, lpos=(0, 0, 0), lquat=(1, 0, 0, 0), surface=None):
        d = (C.c_double * 4)(*([float(x) for x in dims] + [0.0] * (4 - len(dims))))
        p = (C.c_double * 3)(*map(float, lpos))
        q = (C.c_double * 4)(*map(float, lquat))
        s = self._surface(surface)
        return self._all(self.lib.orc_geom_create, cls, body, d, p, q, C.byref(s) if s else None)

    def geom_create_heightfield(self, heights, dx, x0, y0, surface=None):
        hts = np.ascontiguousarray(heights, dtype=np.float32)
        ny, nx = hts.shape
        s = self._surface(surface)
        return self._all(self.lib.orc_geom_create_heightfield, hts.ctypes.data_as(C.c_void_p), nx, ny, dx, x0, y0,
                         C.byref(s) if s else None)

    def collide(self):
        for w in self.worlds:
            self.lib.orc_collide(w)

    def collide_get_pairs(self, world, max_pairs=4096):
        buf = (C.c_int * max_pairs)()
        n = self.lib.orc_collide_get_pairs(self.worlds[world], buf, max_pairs)
        return [(v & 0xffff, v >> 16) for v in buf[:n]]

    # step
    def _post_arrays(self):
        """per-world arrays of the post-step bookkeeping (sized on first use / when the scene grew)"""
        nb = self.lib.orc_world_num_bodies(self.worlds[0])
        nj = self.lib.orc_world_num_joints(self.worlds[0])
        W = self.num_worlds
        if getattr(self, "_post_shape", None) != (nb, nj):
            old_last = getattr(self, "_last", None)
            self._last = np.zeros((W, max(nb, 1) * 6))
            if old_last is not None:
                n = min(old_last.shape[1], self._last.shape[1])
                self._last[:, :n] = old_last[:, :n]
            self._acc = np.zeros((W, max(nb, 1) * 6))
            self._cforce = np.zeros((W, max(nb, 1) * 3))
            self._jupd = np.zeros((W, max(nj, 1) * 7))
            self._post_shape = (nb, nj)
        lin = np.ones(max(nb, 1))
        ang = np.ones(max(nb, 1))
        for b, (l, a) in getattr(self, "_damping", {}).items():
            lin[b], ang[b] = l, a
        fr = -0.1 * np.ones(max(nj, 1))
        for j, k in getattr(self, "_friction", {}).items():
            fr[j] = k
        return lin, ang, fr

    def step(self, h):
        """dWorldQuickStep, then what WorldPhysics::stepTheWorld does after it (joint update + friction, frame
        update, contact force) -- as b2p_step does for the batch"""
        lin, ang, fr = self._post_arrays()
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        for i, w in enumerate(self.worlds):
            self.lib.orc_world_quickstep(w, float(h))
            self.lib.orc_post_step_world(w, float(h), vp(lin), vp(ang), vp(fr), vp(self._last[i]), vp(self._acc[i]),
                                         vp(self._cforce[i]), vp(self._jupd[i]))

    def collide_step(self, h):
        self.collide()
        self.step(h)

    def body_set_damping(self, b, linear, angular):
        self.__dict__.setdefault("_damping", {})[b] = (float(linear), float(angular))

    def joint_set_friction_coefficient(self, j, k):
        self.__dict__.setdefault("_friction", {})[j] = float(k)

    def body_get_acceleration(self, b, world=0):
        a = self._acc[world, 6 * b:6 * b + 6]
        return a[:3].copy(), a[3:].copy()

    def body_get_contact_force(self, b, world=0):
        return self._cforce[world, 3 * b:3 * b + 3].copy()

    def joint_get_update(self, j, world=0):
        return self._jupd[world, 7 * j:7 * j + 7].copy()

    def body_post_download(self, num_bodies):
        """[num_bodies][15][W] like Arena.body_post_download"""
        W = self.num_worlds
        out = np.zeros((num_bodies, 15, W))
        for b in range(num_bodies):
            out[b, 0:6, :] = self._last[:, 6 * b:6 * b + 6].T
            out[b, 6:12, :] = self._acc[:, 6 * b:6 * b + 6].T
            out[b, 12:15, :] = self._cforce[:, 3 * b:3 * b + 3].T
        return out

    def joint_update_download(self, num_joints):
        return np.ascontiguousarray(self._jupd[:, :7 * num_joints].reshape(self.num_worlds, num_joints, 7).transpose(1, 2, 0))

    def sync(self):
        pass

    def bench_run(self, lo, hi, steps, h, do_collide):
        """C loop over worlds[lo:hi] (ctypes releases the GIL, so threads run in parallel)."""
        n = hi - lo
        arr = (C.c_void_p * n)(*self.worlds[lo:hi])
        self.lib.orc_bench_run(arr, n, steps, float(h), int(do_collide))

    def last_num_rows(self, world=0):
        return self.lib.orc_world_last_num_rows(self.worlds[world])

    def last_residual(self, world=0):
        return self.lib.orc_world_last_residual(self.worlds[world])
