"""ctypes binding of the C ABI in include/mars_ode_b200.h.

This module is the only place the shared library is loaded.  There is no fallback: if the
library is missing, import fails with an explicit message (build it with
``python -m mars_ode_physics_b200.build``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
ALL_WORLDS = -1
STATIC_BODY = -1

JOINT_HINGE, JOINT_HINGE2, JOINT_SLIDER, JOINT_FIXED = 1, 2, 3, 6
JOINT_BALL, JOINT_UNIVERSAL = 4, 5
(PARAM_LO_STOP, PARAM_HI_STOP, PARAM_VEL, PARAM_LO_VEL, PARAM_HI_VEL, PARAM_FMAX, PARAM_FUDGE_FACTOR,
 PARAM_BOUNCE, PARAM_CFM, PARAM_STOP_ERP, PARAM_STOP_CFM, PARAM_SUSPENSION_ERP, PARAM_SUSPENSION_CFM,
 PARAM_ERP) = range(14)
PARAM_GROUP2 = 0x100
CONTACT_MU2, CONTACT_SOFT_ERP, CONTACT_SOFT_CFM, CONTACT_ROLLING = 0x001, 0x008, 0x010, 0x400
CONTACT_APPROX1 = 0x7000
REORDER_RANDOMLY, REORDER_NONE = 0, 1
ORDER_WORLD_BLOCK, ORDER_ODE_ISLANDS = 0, 1
SOR_AUTO, SOR_PAIRED, SOR_TMEM = -1, 0, 1
STEPPER_QUICK, STEPPER_DENSE = 0, 1  # dWorldQuickStep | dWorldStep (fast_step == false)
GEOM_SPHERE, GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_PLANE, GEOM_HEIGHTFIELD = range(6)
E_REJECTED = -6


class ContactDesc(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("normal", C.c_double * 3), ("depth", C.c_double),
                ("mode", C.c_int), ("mu", C.c_double), ("mu2", C.c_double), ("rho", C.c_double),
                ("rho2", C.c_double), ("rhoN", C.c_double), ("soft_erp", C.c_double),
                ("soft_cfm", C.c_double)]


class Surface(C.Structure):
    _fields_ = [("mu", C.c_double), ("mu2", C.c_double), ("rho", C.c_double), ("rho2", C.c_double),
                ("rhoN", C.c_double), ("soft_erp", C.c_double), ("soft_cfm", C.c_double)]


def library_path(strict=False):
    name = "libmars_ode_b200_strict.so" if strict else "libmars_ode_b200.so"
    return os.path.join(_HERE, "lib", name)


_P = C.c_void_p
_D = C.c_double
_I = C.c_int
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)
_FP = C.POINTER(C.c_float)

# name -> (restype, argtypes); every symbol declared in include/mars_ode_b200.h
SIGNATURES = {
    "b2p_last_error": (C.c_char_p, []),
    "b2p_version": (_I, []),
    "b2p_device_count": (_I, []),
    "b2p_arena_create": (_I, [_I, _I, C.POINTER(_P)]),
    "b2p_arena_destroy": (None, [_P]),
    "b2p_arena_num_worlds": (_I, [_P]),
    "b2p_arena_set_stream": (_I, [_P, _P]),
    "b2p_arena_set_max_contacts": (_I, [_P, _I]),
    "b2p_arena_enable_rolling_friction": (_I, [_P, _I]),
    "b2p_arena_set_resident_rows": (_I, [_P, _I]),
    "b2p_arena_set_register_rows": (_I, [_P, _I]),
    "b2p_arena_set_sor_kernel": (_I, [_P, _I]),
    "b2p_arena_sor_kernel_used": (_I, [_P]),
    "b2p_arena_set_sor_warps": (_I, [_P, _I]),
    "b2p_arena_set_stepper": (_I, [_P, _I]),
    "b2p_arena_stepper": (_I, [_P]),
    "b2p_dense_failures": (_I, [_P, C.POINTER(C.c_longlong)]),
    "b2p_dense_residual": (_I, [_P, _P]),
    "b2p_world_set_gravity": (_I, [_P, _D, _D, _D]),
    "b2p_world_set_cfm": (_I, [_P, _D]),
    "b2p_world_set_erp": (_I, [_P, _D]),
    "b2p_world_set_max_angular_speed": (_I, [_P, _D]),
    "b2p_world_set_contact_max_correcting_vel": (_I, [_P, _D]),
    "b2p_world_set_contact_surface_layer": (_I, [_P, _D]),
    "b2p_world_set_quickstep_num_iterations": (_I, [_P, _I]),
    "b2p_world_set_quickstep_w": (_I, [_P, _D]),
    "b2p_world_set_reorder_method": (_I, [_P, _I]),
    "b2p_world_rand_set_seed": (_I, [_P, C.c_uint32]),
    "b2p_world_rand_get_seed": (_I, [_P, _I, C.POINTER(C.c_uint32)]),
    "b2p_body_create": (_I, [_P]),
    "b2p_body_set_mass": (_I, [_P, _I, _D, _DP]),
    "b2p_body_set_position": (_I, [_P, _I, _I, _D, _D, _D]),
    "b2p_body_set_quaternion": (_I, [_P, _I, _I, _D, _D, _D, _D]),
    "b2p_body_set_linear_vel": (_I, [_P, _I, _I, _D, _D, _D]),
    "b2p_body_set_angular_vel": (_I, [_P, _I, _I, _D, _D, _D]),
    "b2p_body_set_force": (_I, [_P, _I, _I, _D, _D, _D]),
    "b2p_body_set_torque": (_I, [_P, _I, _I, _D, _D, _D]),
    "b2p_body_add_force": (_I, [_P, _I, _I, _D, _D, _D]),
    "b2p_body_add_torque": (_I, [_P, _I, _I, _D, _D, _D]),
    "b2p_body_add_force_at_pos": (_I, [_P, _I, _I, _D, _D, _D, _D, _D, _D]),
    "b2p_body_set_gravity_mode": (_I, [_P, _I, _I]),
    "b2p_body_set_gyroscopic_mode": (_I, [_P, _I, _I]),
    "b2p_body_get_position": (_I, [_P, _I, _I, _DP]),
    "b2p_body_get_quaternion": (_I, [_P, _I, _I, _DP]),
    "b2p_body_get_rotation": (_I, [_P, _I, _I, _DP]),
    "b2p_body_get_linear_vel": (_I, [_P, _I, _I, _DP]),
    "b2p_body_get_angular_vel": (_I, [_P, _I, _I, _DP]),
    "b2p_body_get_force": (_I, [_P, _I, _I, _DP]),
    "b2p_body_get_torque": (_I, [_P, _I, _I, _DP]),
    "b2p_body_get_state": (_I, [_P, _I, _I, _DP]),
    "b2p_body_set_state": (_I, [_P, _I, _I, _DP]),
    "b2p_state_upload": (_I, [_P, _P]),
    "b2p_state_download": (_I, [_P, _P]),
    "b2p_body_get_state_batch": (_I, [_P, _I, _P]),
    "b2p_joint_obs_download": (_I, [_P, _P]),
    "b2p_joint_feedback_download": (_I, [_P, _P]),
    "b2p_device_joint_feedback": (_P, [_P]),
    "b2p_device_contact_records": (_P, [_P]),
    "b2p_device_contact_count": (_P, [_P]),
    "b2p_body_get_state_batch_begin": (_I, [_P, _I, _P]),
    "b2p_io_wait": (_I, [_P, _I]),
    "b2p_padded_worlds": (_I, [_P]),
    "b2p_device_state": (_P, [_P]),
    "b2p_device_joint_cmd": (_P, [_P]),
    "b2p_device_joint_obs": (_P, [_P]),
    "b2p_joint_create": (_I, [_P, _I, _I, _I]),
    "b2p_joint_set_anchor": (_I, [_P, _I, _D, _D, _D]),
    "b2p_joint_set_axis1": (_I, [_P, _I, _D, _D, _D]),
    "b2p_joint_set_axis2": (_I, [_P, _I, _D, _D, _D]),
    "b2p_joint_set_fixed": (_I, [_P, _I]),
    "b2p_joint_set_param": (_I, [_P, _I, _I, _I, _D]),
    "b2p_joint_get_param": (_I, [_P, _I, _I, _I, _DP]),
    "b2p_joint_set_param_batch": (_I, [_P, _I, _I, _P]),
    "b2p_joints_set_param_batch": (_I, [_P, _I, _P, _I]),
    "b2p_joint_get_anchor": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_get_anchor2": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_get_axis1": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_get_axis2": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_get_angle1": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_get_angle1_rate": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_get_angle2": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_get_angle2_rate": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_add_torque": (_I, [_P, _I, _I, _D]),
    "b2p_joint_get_feedback": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_get_num_rows": (_I, [_P, _I, _I, _IP]),
    "b2p_are_connected_excluding_contacts": (_I, [_P, _I, _I]),
    "b2p_contacts_clear": (_I, [_P, _I]),
    "b2p_contact_add": (_I, [_P, _I, _I, _I, C.POINTER(ContactDesc)]),
    "b2p_contact_count": (_I, [_P, _I, _IP]),
    "b2p_contact_get_feedback": (_I, [_P, _I, _I, _DP]),
    "b2p_contact_get": (_I, [_P, _I, _I, _IP, _IP, C.POINTER(ContactDesc)]),
    "b2p_geom_create": (_I, [_P, _I, _I, _DP, _DP, _DP, C.POINTER(Surface)]),
    "b2p_geom_create_heightfield": (_I, [_P, _P, _I, _I, _D, _D, _D, C.POINTER(Surface)]),
    "b2p_collide": (_I, [_P]),
    "b2p_collide_get_pairs": (_I, [_P, _I, _IP, _I]),
    "b2p_step": (_I, [_P, _D]),
    "b2p_sync": (_I, [_P]),
    "b2p_world_last_num_rows": (_I, [_P, _I, _IP]),
    "b2p_kernel_launch_count": (_I, [_P]),
    "b2p_debug_kernel_times": (_I, [_P, _I, _DP]),
    "b2p_algorithmic_bytes_per_world_step": (_D, [_P]),
    "b2p_arena_set_graph_launch": (_I, [_P, _I]),
    "b2p_arena_set_tune": (_I, [_P, _I]),
    "b2p_arena_set_post_step_outputs": (_I, [_P, _I]),
    "b2p_body_set_damping": (_I, [_P, _I, _D, _D]),
    "b2p_body_get_damping": (_I, [_P, _I, _DP, _DP]),
    "b2p_body_get_acceleration": (_I, [_P, _I, _I, _DP, _DP]),
    "b2p_body_get_contact_force": (_I, [_P, _I, _I, _DP]),
    "b2p_body_post_download": (_I, [_P, _P]),
    "b2p_device_body_post": (_P, [_P]),
    "b2p_observe_begin": (_I, [_P, _I, _P, _P]),
    "b2p_joint_set_friction_coefficient": (_I, [_P, _I, _D]),
    "b2p_joint_get_friction_coefficient": (_I, [_P, _I, _DP]),
    "b2p_joint_get_update": (_I, [_P, _I, _I, _DP]),
    "b2p_joint_update_download": (_I, [_P, _P]),
    "b2p_device_joint_update": (_P, [_P]),
    "b2p_collide_step": (_I, [_P, _D]),
    "b2p_contacts_dropped": (_I, [_P, C.POINTER(C.c_longlong)]),
    "b2p_collide_record_pairs": (_I, [_P, _I]),
    "b2p_debug_lcp_residual": (_I, [_P, _P]),
    "b2p_world_set_row_order": (_I, [_P, _I]),
    "b2p_host_alloc": (_P, [C.c_size_t]),
    "b2p_host_free": (None, [_P]),
}

_libs = {}


def load(strict=False):
    """Load (once) and return the ctypes handle of the CUDA library."""
    if strict in _libs:
        return _libs[strict]
    path = library_path(strict)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the CUDA extension is not built and there is no CPU fallback. "
            "Run `python -m mars_ode_physics_b200.build`.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _libs[strict] = lib
    return lib


class B2PError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"b2p error {code}: {message}")
        self.code = code
