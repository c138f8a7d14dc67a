"""Thin object wrapper over the C ABI (one method per entry point, same argument meaning).

PyTorch is not needed here; bench.py uses torch only for streams/events and pinned memory.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import ALL_WORLDS, B2PError


class Arena:
    """A scene template replicated over ``num_worlds`` independent worlds on one B200."""

    def __init__(self, num_worlds, device=0, strict=False):
        self.lib = capi.load(strict)
        self.h = C.c_void_p()
        self._check(self.lib.b2p_arena_create(int(num_worlds), int(device), C.byref(self.h)))
        self.num_worlds = int(num_worlds)
        self.device = int(device)

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc < 0:
            raise B2PError(rc, self.lib.b2p_last_error().decode())
        return rc

    def close(self):
        if self.h:
            self.lib.b2p_arena_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _out(n):
        return (C.c_double * n)()

    # -- world ------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_handle):
        self._check(self.lib.b2p_arena_set_stream(self.h, C.c_void_p(cuda_stream_handle)))

    def set_max_contacts(self, n):
        self._check(self.lib.b2p_arena_set_max_contacts(self.h, n))

    def enable_rolling_friction(self, on=True):
        self._check(self.lib.b2p_arena_enable_rolling_friction(self.h, int(on)))

    def set_resident_rows(self, n):
        self._check(self.lib.b2p_arena_set_resident_rows(self.h, int(n)))

    def set_sor_kernel(self, which):
        """-1 = automatic, 0 = paired-lane kernel (b2p_sor.cu), 1 = Tensor-Memory kernel (b2p_sor_tmem.cu)"""
        self._check(self.lib.b2p_arena_set_sor_kernel(self.h, int(which)))

    def set_sor_warps(self, warps):
        self._check(self.lib.b2p_arena_set_sor_warps(self.h, int(warps)))

    def sor_kernel_used(self):
        return int(self.lib.b2p_arena_sor_kernel_used(self.h))

    def set_register_rows(self, n):
        self._check(self.lib.b2p_arena_set_register_rows(self.h, int(n)))

    def set_gravity(self, x, y, z):
        self._check(self.lib.b2p_world_set_gravity(self.h, x, y, z))

    def set_cfm(self, v):
        self._check(self.lib.b2p_world_set_cfm(self.h, v))

    def set_erp(self, v):
        self._check(self.lib.b2p_world_set_erp(self.h, v))

    def set_max_angular_speed(self, v):
        self._check(self.lib.b2p_world_set_max_angular_speed(self.h, v))

    def set_contact_max_correcting_vel(self, v):
        self._check(self.lib.b2p_world_set_contact_max_correcting_vel(self.h, v))

    def set_contact_surface_layer(self, v):
        self._check(self.lib.b2p_world_set_contact_surface_layer(self.h, v))

    def set_quickstep_num_iterations(self, n):
        self._check(self.lib.b2p_world_set_quickstep_num_iterations(self.h, n))

    def set_quickstep_w(self, v):
        self._check(self.lib.b2p_world_set_quickstep_w(self.h, v))

    def set_reorder_method(self, m):
        self._check(self.lib.b2p_world_set_reorder_method(self.h, m))

    def set_order_mode(self, mode):
        """0: one block per world (default); 1: ODE's island order (capi.ORDER_ODE_ISLANDS)"""
        self._check(self.lib.b2p_world_set_row_order(self.h, int(mode)))

    def rand_set_seed(self, seed):
        self._check(self.lib.b2p_world_rand_set_seed(self.h, seed))

    def rand_get_seed(self, world):
        s = C.c_uint32()
        self._check(self.lib.b2p_world_rand_get_seed(self.h, world, C.byref(s)))
        return s.value

    # -- bodies -----------------------------------------------------------------------------
    def body_create(self):
        return self._check(self.lib.b2p_body_create(self.h))

    def body_set_mass(self, b, mass, I):
        arr = (C.c_double * 9)(*[float(x) for x in np.asarray(I, dtype=float).reshape(9)])
        self._check(self.lib.b2p_body_set_mass(self.h, b, mass, arr))

    def body_set_position(self, b, p, world=ALL_WORLDS):
        self._check(self.lib.b2p_body_set_position(self.h, b, world, *map(float, p)))

    def body_set_quaternion(self, b, q, world=ALL_WORLDS):
        self._check(self.lib.b2p_body_set_quaternion(self.h, b, world, *map(float, q)))

    def body_set_linear_vel(self, b, v, world=ALL_WORLDS):
        self._check(self.lib.b2p_body_set_linear_vel(self.h, b, world, *map(float, v)))

    def body_set_angular_vel(self, b, v, world=ALL_WORLDS):
        self._check(self.lib.b2p_body_set_angular_vel(self.h, b, world, *map(float, v)))

    def body_set_force(self, b, v, world=ALL_WORLDS):
        self._check(self.lib.b2p_body_set_force(self.h, b, world, *map(float, v)))

    def body_set_torque(self, b, v, world=ALL_WORLDS):
        self._check(self.lib.b2p_body_set_torque(self.h, b, world, *map(float, v)))

    def body_add_force(self, b, v, world=ALL_WORLDS):
        self._check(self.lib.b2p_body_add_force(self.h, b, world, *map(float, v)))

    def body_add_torque(self, b, v, world=ALL_WORLDS):
        self._check(self.lib.b2p_body_add_torque(self.h, b, world, *map(float, v)))

    def body_add_force_at_pos(self, b, f, p, world=ALL_WORLDS):
        self._check(self.lib.b2p_body_add_force_at_pos(self.h, b, world, *map(float, f), *map(float, p)))

    def body_set_gravity_mode(self, b, on):
        self._check(self.lib.b2p_body_set_gravity_mode(self.h, b, int(on)))

    def body_set_gyroscopic_mode(self, b, on):
        self._check(self.lib.b2p_body_set_gyroscopic_mode(self.h, b, int(on)))

    def _get(self, fn, n, *ids):
        out = self._out(n)
        self._check(fn(self.h, *ids, out))
        return np.array(out[:], dtype=float)

    def body_get_position(self, b, world=0):
        return self._get(self.lib.b2p_body_get_position, 3, b, world)

    def body_get_quaternion(self, b, world=0):
        return self._get(self.lib.b2p_body_get_quaternion, 4, b, world)

    def body_get_rotation(self, b, world=0):
        return self._get(self.lib.b2p_body_get_rotation, 9, b, world)

    def body_get_linear_vel(self, b, world=0):
        return self._get(self.lib.b2p_body_get_linear_vel, 3, b, world)

    def body_get_angular_vel(self, b, world=0):
        return self._get(self.lib.b2p_body_get_angular_vel, 3, b, world)

    def body_get_force(self, b, world=0):
        return self._get(self.lib.b2p_body_get_force, 3, b, world)

    def body_get_torque(self, b, world=0):
        return self._get(self.lib.b2p_body_get_torque, 3, b, world)

    def body_get_state(self, b, world=0):
        return self._get(self.lib.b2p_body_get_state, 13, b, world)

    def body_set_state(self, b, s, world=ALL_WORLDS):
        arr = (C.c_double * 13)(*map(float, s))
        self._check(self.lib.b2p_body_set_state(self.h, b, world, arr))

    def state_upload(self, host_state):
        a = np.ascontiguousarray(host_state, dtype=np.float32)
        self._check(self.lib.b2p_state_upload(self.h, a.ctypes.data_as(C.c_void_p)))

    def state_download(self, num_bodies, out=None):
        if out is None:
            out = np.empty((num_bodies, 13, self.num_worlds), dtype=np.float32)
        self._check(self.lib.b2p_state_download(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def joint_obs_download(self, num_joints):
        """float[num_joints][4][num_worlds]: angle1, rate1, angle2, rate2 written by the last step."""
        out = np.empty((num_joints, 4, self.num_worlds), dtype=np.float32)
        self._check(self.lib.b2p_joint_obs_download(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def joint_feedback_download(self, num_joints):
        """float[num_joints][13][num_worlds]: f1 t1 f2 t2 lambda of every joint after the last step."""
        out = np.empty((num_joints, 13, self.num_worlds), dtype=np.float32)
        self._check(self.lib.b2p_joint_feedback_download(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def body_get_state_batch_ptr(self, b, host_ptr):
        """Asynchronous D2H of float[13][num_worlds] into (pinned) host memory; sync() before reading."""
        self._check(self.lib.b2p_body_get_state_batch(self.h, b, C.c_void_p(host_ptr)))

    def body_get_state_batch_begin(self, b, host_ptr):
        """Pipelined D2H (copy stream) of float[13][num_worlds]; returns a ticket for io_wait()."""
        return self._check(self.lib.b2p_body_get_state_batch_begin(self.h, b, C.c_void_p(host_ptr)))

    def io_wait(self, ticket):
        self._check(self.lib.b2p_io_wait(self.h, int(ticket)))

    # -- joints -----------------------------------------------------------------------------
    def joint_create(self, jtype, b1, b2):
        return self._check(self.lib.b2p_joint_create(self.h, jtype, b1, b2))

    def joint_set_anchor(self, j, p):
        self._check(self.lib.b2p_joint_set_anchor(self.h, j, *map(float, p)))

    def joint_set_axis1(self, j, a):
        self._check(self.lib.b2p_joint_set_axis1(self.h, j, *map(float, a)))

    def joint_set_axis2(self, j, a):
        self._check(self.lib.b2p_joint_set_axis2(self.h, j, *map(float, a)))

    def joint_set_fixed(self, j):
        self._check(self.lib.b2p_joint_set_fixed(self.h, j))

    def joint_set_param(self, j, param, value, world=ALL_WORLDS):
        self._check(self.lib.b2p_joint_set_param(self.h, j, world, param, float(value)))

    def joint_get_param(self, j, param, world=0):
        v = C.c_double()
        self._check(self.lib.b2p_joint_get_param(self.h, j, world, param, C.byref(v)))
        return v.value

    def joint_set_param_batch(self, j, param, values):
        a = np.ascontiguousarray(values, dtype=np.float32)
        assert a.shape == (self.num_worlds,)
        self._check(self.lib.b2p_joint_set_param_batch(self.h, j, param, a.ctypes.data_as(C.c_void_p)))

    def joints_set_param_batch_ptr(self, param, host_ptr, num_joints):
        """One H2D for values[num_joints][num_worlds] (pinned host memory) of joints 0 .. num_joints-1."""
        self._check(self.lib.b2p_joints_set_param_batch(self.h, param, C.c_void_p(host_ptr), int(num_joints)))

    def joint_set_param_batch_ptr(self, j, param, host_ptr):
        self._check(self.lib.b2p_joint_set_param_batch(self.h, j, param, C.c_void_p(host_ptr)))

    def joint_get_anchor(self, j, world=0):
        return self._get(self.lib.b2p_joint_get_anchor, 3, j, world)

    def joint_get_anchor2(self, j, world=0):
        return self._get(self.lib.b2p_joint_get_anchor2, 3, j, world)

    def joint_get_axis1(self, j, world=0):
        return self._get(self.lib.b2p_joint_get_axis1, 3, j, world)

    def joint_get_axis2(self, j, world=0):
        return self._get(self.lib.b2p_joint_get_axis2, 3, j, world)

    def _scalar(self, fn, j, world):
        v = C.c_double()
        self._check(fn(self.h, j, world, C.byref(v)))
        return v.value

    def joint_get_angle1(self, j, world=0):
        return self._scalar(self.lib.b2p_joint_get_angle1, j, world)

    def joint_get_angle1_rate(self, j, world=0):
        return self._scalar(self.lib.b2p_joint_get_angle1_rate, j, world)

    def joint_get_angle2(self, j, world=0):
        return self._scalar(self.lib.b2p_joint_get_angle2, j, world)

    def joint_get_angle2_rate(self, j, world=0):
        return self._scalar(self.lib.b2p_joint_get_angle2_rate, j, world)

    def joint_add_torque(self, j, t, world=ALL_WORLDS):
        self._check(self.lib.b2p_joint_add_torque(self.h, j, world, float(t)))

    def joint_get_feedback(self, j, world=0):
        return self._get(self.lib.b2p_joint_get_feedback, 13, j, world)

    def joint_get_num_rows(self, j, world=0):
        v = C.c_int()
        self._check(self.lib.b2p_joint_get_num_rows(self.h, j, world, C.byref(v)))
        return v.value

    # -- contacts ---------------------------------------------------------------------------
    def contacts_clear(self, world=ALL_WORLDS):
        self._check(self.lib.b2p_contacts_clear(self.h, world))

    def contact_add(self, world, b1, b2, **kw):
        """Returns the contact id, or -1 when the bodies are joint-connected (rejected)."""
        d = capi.ContactDesc()
        d.pos[:] = [float(x) for x in kw["pos"]]
        d.normal[:] = [float(x) for x in kw["normal"]]
        d.depth = float(kw["depth"])
        d.mode = int(kw["mode"])
        for k in ("mu", "mu2", "rho", "rho2", "rhoN", "soft_erp", "soft_cfm"):
            setattr(d, k, float(kw.get(k, 0.0)))
        rc = self.lib.b2p_contact_add(self.h, world, b1, b2, C.byref(d))
        if rc == capi.E_REJECTED:
            return -1
        return self._check(rc)

    def contact_count(self, world=0):
        v = C.c_int()
        self._check(self.lib.b2p_contact_count(self.h, world, C.byref(v)))
        return v.value

    def contact_get_feedback(self, world, c):
        return self._get(self.lib.b2p_contact_get_feedback, 3, world, c)

    def contact_get(self, world, c):
        b1, b2, d = C.c_int(), C.c_int(), capi.ContactDesc()
        self._check(self.lib.b2p_contact_get(self.h, world, c, C.byref(b1), C.byref(b2), C.byref(d)))
        return b1.value, b2.value, d

    # -- geoms / collision ------------------------------------------------------------------
    @staticmethod
    def _surface(s):
        if s is None:
            return None
        out = capi.Surface()
        for k in ("mu", "mu2", "rho", "rho2", "rhoN", "soft_erp", "soft_cfm"):
            setattr(out, k, float(s.get(k, 0.0)))
        return out

    def geom_create(self, cls, body, dims, lpos=(0, 0, 0), lquat=(1, 0, 0, 0), surface=None):
        d = (C.c_double * 4)(*([float(x) for x in dims] + [0.0] * (4 - len(dims))))
        p = (C.c_double * 3)(*map(float, lpos))
        q = (C.c_double * 4)(*map(float, lquat))
        s = self._surface(surface)
        return self._check(self.lib.b2p_geom_create(self.h, cls, body, d, p, q, C.byref(s) if s else None))

    def geom_create_heightfield(self, heights, dx, x0, y0, surface=None):
        hts = np.ascontiguousarray(heights, dtype=np.float32)
        ny, nx = hts.shape
        s = self._surface(surface)
        return self._check(self.lib.b2p_geom_create_heightfield(
            self.h, hts.ctypes.data_as(C.c_void_p), nx, ny, dx, x0, y0, C.byref(s) if s else None))

    def collide(self):
        self._check(self.lib.b2p_collide(self.h))

    def collide_get_pairs(self, world, max_pairs=4096):
        buf = (C.c_int * max_pairs)()
        n = self._check(self.lib.b2p_collide_get_pairs(self.h, world, buf, max_pairs))
        return [(v & 0xffff, v >> 16) for v in buf[:n]]

    # -- step -------------------------------------------------------------------------------
    def step(self, h):
        self._check(self.lib.b2p_step(self.h, float(h)))

    def sync(self):
        self._check(self.lib.b2p_sync(self.h))

    def last_num_rows(self, world=0):
        v = C.c_int()
        self._check(self.lib.b2p_world_last_num_rows(self.h, world, C.byref(v)))
        return v.value

    def kernel_launch_count(self):
        return self.lib.b2p_kernel_launch_count(self.h)

    def debug_kernel_times(self, enable=True):
        """mean ms of the (assemble, SOR, integrate) kernels since timing was enabled, steps counted"""
        out = self._out(4)
        self._check(self.lib.b2p_debug_kernel_times(self.h, int(enable), out))
        return np.array(out[:], dtype=float)

    def algorithmic_bytes_per_world_step(self):
        return self.lib.b2p_algorithmic_bytes_per_world_step(self.h)

    def padded_worlds(self):
        return self.lib.b2p_padded_worlds(self.h)

    # -- zero-copy views of the batched device arrays (CUDA array interface / DLPack) ---------------
    _DEVICE_ARRAYS = {  # name -> (C ABI accessor, fields per item, which count, dtype)
        "state": ("b2p_device_state", 13, "bodies", "<f4"),                   # pos3 quat4 lvel3 avel3
        "joint_cmd": ("b2p_device_joint_cmd", 4, "joints", "<f4"),            # vel1 fmax1 vel2 fmax2
        "joint_obs": ("b2p_device_joint_obs", 4, "joints", "<f4"),            # angle1 rate1 angle2 rate2
        "joint_feedback": ("b2p_device_joint_feedback", 13, "joints", "<f4"), # f1 t1 f2 t2 lambda
        "joint_update": ("b2p_device_joint_update", 7, "joints", "<f4"),      # motor torque, axis torque3, load3
        "body_post": ("b2p_device_body_post", 15, "bodies", "<f4"),           # last vel6, acc6, contact force3
        "contact_records": ("b2p_device_contact_records", 16, "contacts", "<f4"),
        "contact_count": ("b2p_device_contact_count", 0, None, "<i4"),
    }

    def device_array(self, name, count=None):
        """The device array ``name`` of the last step as an object with ``__cuda_array_interface__`` (version 3) and
        ``__dlpack__``: shape ``[count, fields, padded_worlds]`` (world index fastest, worlds padded to a multiple of
        32), float32, no copy.  ``count`` = number of bodies / joints / contact slots (the arrays hold all of the
        scene's; pass fewer to view a prefix).  ``torch.as_tensor(arena.device_array("joint_obs", nj), device="cuda")``
        or ``torch.from_dlpack(...)`` give a tensor on the arena's GPU; the arena keeps ownership, and the content is
        defined once the arena's stream has reached the end of the step (``sync()`` or stream ordering)."""
        fn, fields, _, typestr = self._DEVICE_ARRAYS[name]
        ptr = getattr(self.lib, fn)(self.h)
        if not ptr:
            raise B2PError(-1, f"device array {name!r} does not exist (yet)")
        if fields and count is None:
            raise ValueError("count (bodies / joints / contact slots) is required")
        Wp = self.padded_worlds()
        shape = (Wp,) if fields == 0 else (int(count), fields, Wp)
        return _DeviceArray(ptr, shape, typestr, self.device)

    # -- round 2: graph launch, post-step outputs, diagnostics ------------------------------------
    def set_tune(self, bits):
        self._check(self.lib.b2p_arena_set_tune(self.h, int(bits)))

    def set_graph_launch(self, on=True):
        self._check(self.lib.b2p_arena_set_graph_launch(self.h, int(on)))

    def set_post_step_outputs(self, flags):
        self._check(self.lib.b2p_arena_set_post_step_outputs(self.h, int(flags)))

    def collide_step(self, h):
        """collide() + step(h) as one call (one CUDA graph launch once the parameters are stable)"""
        self._check(self.lib.b2p_collide_step(self.h, float(h)))

    def collide_record_pairs(self, on=True):
        self._check(self.lib.b2p_collide_record_pairs(self.h, int(on)))

    def contacts_dropped(self):
        v = C.c_longlong()
        self._check(self.lib.b2p_contacts_dropped(self.h, C.byref(v)))
        return v.value

    def body_set_damping(self, b, linear, angular):
        self._check(self.lib.b2p_body_set_damping(self.h, b, float(linear), float(angular)))

    def body_get_damping(self, b):
        l, a = C.c_double(), C.c_double()
        self._check(self.lib.b2p_body_get_damping(self.h, b, C.byref(l), C.byref(a)))
        return l.value, a.value

    def body_get_acceleration(self, b, world=0):
        l, a = self._out(3), self._out(3)
        self._check(self.lib.b2p_body_get_acceleration(self.h, b, world, l, a))
        return np.array(l[:], dtype=float), np.array(a[:], dtype=float)

    def body_get_contact_force(self, b, world=0):
        return self._get(self.lib.b2p_body_get_contact_force, 3, b, world)

    def body_post_download(self, num_bodies):
        """float[num_bodies][15][num_worlds]: last_l_vel3 last_a_vel3 l_acc3 a_acc3 contact_force3"""
        out = np.empty((num_bodies, 15, self.num_worlds), dtype=np.float32)
        self._check(self.lib.b2p_body_post_download(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def joint_set_friction_coefficient(self, j, k):
        self._check(self.lib.b2p_joint_set_friction_coefficient(self.h, j, float(k)))

    def joint_get_friction_coefficient(self, j):
        v = C.c_double()
        self._check(self.lib.b2p_joint_get_friction_coefficient(self.h, j, C.byref(v)))
        return v.value

    def joint_get_update(self, j, world=0):
        """motor_torque, axis1_torque[3], joint_load[3] of the last step (Joint::update)"""
        return self._get(self.lib.b2p_joint_get_update, 7, j, world)

    def joint_update_download(self, num_joints):
        out = np.empty((num_joints, 7, self.num_worlds), dtype=np.float32)
        self._check(self.lib.b2p_joint_update_download(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def observe_begin(self, b, host_body_ptr, host_jobs_ptr=None):
        """Pipelined D2H of body b's state [13][W] and (optionally) all joint observations [nj][4][W]."""
        return self._check(self.lib.b2p_observe_begin(self.h, b, C.c_void_p(host_body_ptr),
                                                      C.c_void_p(host_jobs_ptr) if host_jobs_ptr else None))

    def set_stepper(self, stepper):
        """capi.STEPPER_QUICK (dWorldQuickStep) or capi.STEPPER_DENSE (dWorldStep: exact dense LCP)"""
        self._check(self.lib.b2p_arena_set_stepper(self.h, int(stepper)))

    def stepper(self):
        return self.lib.b2p_arena_stepper(self.h)

    def dense_failures(self):
        v = C.c_longlong()
        self._check(self.lib.b2p_dense_failures(self.h, C.byref(v)))
        return v.value

    def dense_residual(self):
        """dense stepper: per-world complementarity residual of the last step's solved (scaled) problem"""
        out = np.empty((self.num_worlds,), dtype=np.float32)
        self._check(self.lib.b2p_dense_residual(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def lcp_residual(self):
        """per-world LCP residual of the last step (projected |J invM J^T lambda + cfm lambda - rhs|)"""
        out = np.empty((self.num_worlds,), dtype=np.float32)
        self._check(self.lib.b2p_debug_lcp_residual(self.h, out.ctypes.data_as(C.c_void_p)))
        return out


class _DeviceArray:
    """A borrowed view of device memory: CUDA array interface v3, DLPack through torch when it is importable."""

    def __init__(self, ptr, shape, typestr, device):
        self.device = device
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None, "stream": None}

    def _torch(self):
        import torch
        return torch.as_tensor(self, device=f"cuda:{self.device}")

    def __dlpack__(self, stream=None):
        return self._torch().__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return (2, self.device)  # kDLCUDA
