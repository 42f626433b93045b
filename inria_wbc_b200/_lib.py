"""ctypes binding of include/iwbc_batch.h (libiwbc_b200.so)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ETC_PATH = os.path.join(_HERE, "etc")
ROBOTS_PATH = os.path.join(_HERE, "robots")
STATUS_OPTIMAL = 0
ACTIVE_WORDS = 12


class IwbcError(RuntimeError):
    pass


def lib_path():
    # (IWBC_B200_LIB: another build of the same library, for A/B measurements of compile-time variants)
    return os.environ.get("IWBC_B200_LIB") or os.path.join(_HERE, "libiwbc_b200.so")


_lib = None


class _Results(C.Structure):
    _fields_ = [("tau", C.c_void_p), ("dv", C.c_void_p), ("f", C.c_void_p), ("q_next", C.c_void_p),
                ("v_next", C.c_void_p), ("q_cmd", C.c_void_p), ("momentum", C.c_void_p), ("status", C.c_void_p),
                ("iters", C.c_void_p), ("active", C.c_void_p)]


class _Safety(C.Structure):
    _fields_ = [("send_cmd", C.c_void_p), ("colliding", C.c_void_p), ("q_prev", C.c_void_p), ("v_prev", C.c_void_p), ("q_solver", C.c_void_p),
                ("q_cmd", C.c_void_p), ("tau_cmd", C.c_void_p), ("dq_cmd", C.c_void_p), ("ddq_cmd", C.c_void_p)]


def lib():
    """Load the product library; fails loudly if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise IwbcError("libiwbc_b200.so is not built: run `python __graft_entry__.py` (or make -C inria_wbc_b200/csrc)")
    L = C.CDLL(p)
    vp, cp, ci, cu, cd, cl = C.c_void_p, C.c_char_p, C.c_int, C.c_uint, C.c_double, C.c_long
    pi = C.POINTER(C.c_int)
    sig = {
        "iwbc_last_error": (cp, []),
        "iwbc_abi_version": (ci, []),
        "iwbc_problem_load": (vp, [cp, cp, cp, cp]),
        "iwbc_problem_free": (None, [vp]),
        "iwbc_problem_data": (vp, [vp]),
        "iwbc_problem_dims": (ci, [vp, pi, pi, pi, pi, pi]),
        "iwbc_problem_qp_shape": (ci, [vp, cu, pi, pi, pi, pi]),
        "iwbc_problem_default_contact_mask": (cu, [vp]),
        "iwbc_problem_default_refs": (ci, [vp, vp]),
        "iwbc_problem_ref_slot": (ci, [vp, cp, pi, pi]),
        "iwbc_problem_set_task_weight": (ci, [vp, cp, cd]),
        "iwbc_problem_contact_index": (ci, [vp, cp]),
        "iwbc_problem_name": (ci, [vp, ci, ci, cp, ci]),
        "iwbc_problem_q0": (ci, [vp, vp]),
        "iwbc_problem_closed_loop": (ci, [vp]),
        "iwbc_batch_create": (vp, [vp, ci, ci]),
        "iwbc_batch_free": (None, [vp]),
        "iwbc_batch_update": (ci, [vp, ci, vp, vp, vp, cl, vp, C.POINTER(_Results)]),
        "iwbc_batch_update_device": (ci, [vp, ci, vp, vp, vp, cl, cu, C.POINTER(_Results), vp]),
        "iwbc_batch_rollout_device": (ci, [vp, ci, vp, vp, vp, cl, cu, ci, C.POINTER(_Results), vp]),
        "iwbc_batch_launch_count": (cl, [vp]),
        "iwbc_batch_kernel_ms": (cd, [vp, ci]),
        "iwbc_batch_enable_timing": (ci, [vp, ci]),
        "iwbc_batch_debug_instance": (ci, [vp, vp, vp, vp, cu, vp, vp, vp, vp, vp]),
        "iwbc_batch_task_cost": (ci, [vp, vp, vp, vp, cu, C.c_char_p, vp, vp, vp]),
        "iwbc_batch_objective_value": (ci, [vp, vp, vp, vp, cu, vp]),
        "iwbc_measure_fp64_tflops": (cd, [ci]),
        "iwbc_batch_debug_cycles": (ci, [vp, vp]),
        "iwbc_batch_phase_profile": (ci, [vp, ci, vp]),
        "iwbc_batch_fast_info": (ci, [vp, vp]),
        "iwbc_batch_loop_profile": (ci, [vp, vp]),
        "iwbc_batch_launch_info": (ci, [vp, pi, pi, pi, C.POINTER(C.c_long)]),
        "iwbc_qp_solve_batch": (ci, [ci, ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "iwbc_batch_behavior_run": (ci, [vp, vp, ci, vp, vp, vp, ci, ci, C.POINTER(_Results), vp]),
        "iwbc_batch_behavior_run_trace": (ci, [vp, vp, ci, vp, vp, vp, ci, ci, C.POINTER(_Results), vp, vp]),
        "iwbc_batch_behavior_rollout_trace_device": (ci, [vp, vp, ci, vp, vp, vp, vp, ci, ci, ci, C.POINTER(_Results), vp, vp, vp]),
        "iwbc_collision_load": (vp, [vp, cp]),
        "iwbc_collision_free": (None, [vp]),
        "iwbc_collision_info": (ci, [vp, pi, pi, pi]),
        "iwbc_collision_spheres": (ci, [vp, vp, vp]),
        "iwbc_batch_check_collisions_device": (ci, [vp, vp, ci, vp, vp, vp, vp]),
        "iwbc_batch_safe_rollout_device": (ci, [vp, vp, ci, vp, vp, vp, cl, cu, ci, C.POINTER(_Safety), vp, vp]),
        "iwbc_batch_qp_step_back_device": (ci, [vp, ci, vp, vp, C.POINTER(_Safety), vp, vp]),
        "iwbc_stabilizer_create": (vp, [vp]),
        "iwbc_stabilizer_free": (None, [vp]),
        "iwbc_stabilizer_reset": (ci, [vp, vp]),
        "iwbc_batch_update_sensors_device": (ci, [vp, vp, ci, vp, vp, vp, cl, vp, vp, ci, vp, C.POINTER(_Results), vp, vp]),
        "iwbc_batch_update_sensors": (ci, [vp, vp, ci, vp, vp, vp, cl, vp, vp, ci, vp, C.POINTER(_Results), vp]),
        "iwbc_multi_create": (vp, [vp, ci, vp, ci]),
        "iwbc_multi_free": (None, [vp]),
        "iwbc_multi_device_count": (ci, [vp]),
        "iwbc_multi_update": (ci, [vp, ci, vp, vp, vp, cl, vp, C.POINTER(_Results)]),
        "iwbc_multi_launch_count": (cl, [vp]),
        "iwbc_shard_range": (ci, [ci, ci, ci, pi, pi]),
        "iwbc_host_alloc": (vp, [C.c_size_t]),
        "iwbc_host_free": (None, [vp]),
        "iwbc_behavior_load": (vp, [vp, cp, cp]),
        "iwbc_behavior_free": (None, [vp]),
        "iwbc_behavior_info": (ci, [vp, pi, pi, pi, pi, vp]),
        "iwbc_behavior_segment": (ci, [vp, ci, pi, pi, vp, vp, vp]),
        "iwbc_behavior_eval_device": (ci, [vp, ci, ci, vp, ci, vp, vp, ci, vp]),
        "iwbc_batch_behavior_rollout_device": (ci, [vp, vp, ci, vp, vp, vp, vp, ci, ci, ci, C.POINTER(_Results), vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    L._declared = sorted(sig)
    _lib = L
    return L


def _err(L):
    return (L.iwbc_last_error() or b"").decode()


def _ptr(a):
    return None if a is None else a.ctypes.data


class Problem:
    """A compiled controller (model + task stack): what `controllers::Factory::create(name, yaml)` builds."""

    def __init__(self, controller_yaml, robots_path=ROBOTS_PATH, etc_path=ETC_PATH, overrides=None):
        L = lib()
        if not os.path.isabs(controller_yaml) and not os.path.exists(controller_yaml):
            controller_yaml = os.path.join(ETC_PATH, controller_yaml)
        ov = None
        if overrides:
            ov = overrides if isinstance(overrides, str) else "\n".join("%s: %s" % (k, _yaml_scalar(v)) for k, v in overrides.items())
        self._h = L.iwbc_problem_load(controller_yaml.encode(), robots_path.encode(), etc_path.encode() if etc_path else None,
                                      ov.encode() if ov else None)
        if not self._h:
            raise IwbcError(_err(L))
        d = [C.c_int() for _ in range(5)]
        L.iwbc_problem_dims(self._h, *[C.byref(x) for x in d])
        self.nq, self.nv, self.na, self.ncontacts, self.nref = [x.value for x in d]
        self.default_contact_mask = L.iwbc_problem_default_contact_mask(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.iwbc_problem_free(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    @property
    def data_ptr(self):
        """const iwbc_problem_t* (handed to the CPU oracle by the tests)"""
        return lib().iwbc_problem_data(self._h)

    def struct(self):
        """ctypes view of the flat iwbc_problem_t (read-only)."""
        from ._problem_struct import view
        return view(self.data_ptr)

    def qp_shape(self, contact_mask=None):
        m = self.default_contact_mask if contact_mask is None else contact_mask
        d = [C.c_int() for _ in range(4)]
        lib().iwbc_problem_qp_shape(self._h, m, *[C.byref(x) for x in d])
        return tuple(x.value for x in d)  # n, me, mi, cost rows

    def default_refs(self):
        out = np.zeros(self.nref)
        lib().iwbc_problem_default_refs(self._h, out.ctypes.data)
        return out

    def q0(self):
        out = np.zeros(self.nq)
        lib().iwbc_problem_q0(self._h, out.ctypes.data)
        return out

    def ref_slot(self, name):
        o, n = C.c_int(), C.c_int()
        if lib().iwbc_problem_ref_slot(self._h, name.encode(), C.byref(o), C.byref(n)) != 0:
            raise IwbcError(_err(lib()))
        return o.value, n.value

    def set_task_weight(self, name, w):
        if lib().iwbc_problem_set_task_weight(self._h, name.encode(), float(w)) != 0:
            raise IwbcError(_err(lib()))

    def contact_index(self, name):
        r = lib().iwbc_problem_contact_index(self._h, name.encode())
        if r < 0:
            raise IwbcError(_err(lib()))
        return r

    def names(self, kind):
        k = {"dofs": 0, "tasks": 1, "contacts": 2, "mimics": 3}[kind]
        n = lib().iwbc_problem_name(self._h, k, -1, None, 0)
        out = []
        buf = C.create_string_buffer(256)
        for i in range(n):
            lib().iwbc_problem_name(self._h, k, i, buf, 256)
            out.append(buf.value.decode())
        return out


class CollisionModel:
    """Sphere cover of the robot's links for the post-solve model collision check (utils::CollisionCheck)."""

    def __init__(self, problem, collision_yaml):
        self.problem = problem
        if not os.path.isabs(collision_yaml):
            collision_yaml = os.path.join(ETC_PATH, collision_yaml)
        self._h = lib().iwbc_collision_load(problem.handle, collision_yaml.encode())
        if not self._h:
            raise IwbcError(_err(lib()))
        ns, nm, sk = C.c_int(), C.c_int(), C.c_int()
        lib().iwbc_collision_info(self._h, C.byref(ns), C.byref(nm), C.byref(sk))
        self.nspheres, self.nmembers, self.skipped_links = ns.value, nm.value, sk.value

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.iwbc_collision_free(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def spheres(self):
        """(p_radius[nspheres][4], joint_member[nspheres][2])"""
        pr = np.zeros((self.nspheres, 4))
        jm = np.zeros((self.nspheres, 2), dtype=np.int32)
        lib().iwbc_collision_spheres(self._h, pr.ctypes.data, jm.ctypes.data)
        return pr, jm


class Behavior:
    """A behaviour (reference generator) compiled into a device-evaluable segment table: `humanoid::walk`."""

    def __init__(self, problem, behavior_yaml, overrides=None):
        self.problem = problem
        if not os.path.isabs(behavior_yaml):
            behavior_yaml = os.path.join(ETC_PATH, behavior_yaml)
        ov = None
        if overrides:
            ov = "\n".join("%s: %s" % (k, _yaml_scalar(v)) for k, v in overrides.items()).encode()
        self._h = lib().iwbc_behavior_load(problem.handle, behavior_yaml.encode(), ov)
        if not self._h:
            raise IwbcError(_err(lib()))
        tt, ns, nc, nm = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        masks = (C.c_uint * 8)()
        lib().iwbc_behavior_info(self._h, C.byref(tt), C.byref(ns), C.byref(nc), C.byref(nm), masks)
        self.total_ticks, self.nsegments, self.nchannels = tt.value, ns.value, nc.value
        self.masks = [int(masks[k]) for k in range(nm.value)]

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.iwbc_behavior_free(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def segment(self, k):
        """(start, ticks, mask, mask_last, waypoints[nchannels][8] = p0(3) p1(3) duration minjerk)"""
        st, tk, m, ml = C.c_int(), C.c_int(), C.c_uint(), C.c_uint()
        wp = np.zeros((self.nchannels, 8))
        if lib().iwbc_behavior_segment(self._h, k, C.byref(st), C.byref(tk), C.byref(m), C.byref(ml), wp.ctypes.data) != 0:
            raise IwbcError(_err(lib()))
        return st.value, tk.value, m.value, ml.value, wp

    def eval_device(self, device, B, phase_ptr, tick, refs_ptr, mask_ptr=None, init_refs=True, stream=0):
        rc = lib().iwbc_behavior_eval_device(self._h, device, B, phase_ptr, tick, refs_ptr, mask_ptr, 1 if init_refs else 0, stream)
        if rc != 0:
            raise IwbcError(_err(lib()))


class Stabilizer:
    def __init__(self, batch):
        self.batch = batch
        self._h = lib().iwbc_stabilizer_create(batch._h)
        if not self._h:
            raise IwbcError(_err(lib()))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.iwbc_stabilizer_free(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def reset(self):
        if lib().iwbc_stabilizer_reset(self._h, None) != 0:
            raise IwbcError(_err(lib()))


def _yaml_scalar(v):
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, dict):
        return "{" + ", ".join("%s: %s" % (k, _yaml_scalar(x)) for k, x in v.items()) + "}"
    return str(v)


class Results:
    """Host-side result arrays of one batched update (numpy, instance-major)."""

    FIELDS = ("tau", "dv", "f", "q_next", "v_next", "q_cmd", "momentum", "status", "iters", "active")

    def __init__(self, problem, B, fields=FIELDS):
        P = problem
        shapes = {"tau": (B, P.na), "dv": (B, P.nv), "f": (B, 12 * P.ncontacts), "q_next": (B, P.nq), "v_next": (B, P.nv),
                  "q_cmd": (B, P.nv), "momentum": (B, 3), "status": (B,), "iters": (B,), "active": (B, ACTIVE_WORDS)}
        dt = {"status": np.int32, "iters": np.int32, "active": np.uint32}
        for f in self.FIELDS:
            setattr(self, f, np.zeros(shapes[f], dtype=dt.get(f, np.float64)) if f in fields else None)

    def as_struct(self):
        return _Results(*[_ptr(getattr(self, f)) for f in self.FIELDS])


class Batch:
    """One GPU's batch workspace: `update(states[])` (BASELINE.json north_star)."""

    def __init__(self, problem, device=0, max_batch=65536):
        self.problem = problem
        self._h = lib().iwbc_batch_create(problem.handle, device, max_batch)
        if not self._h:
            raise IwbcError(_err(lib()))
        self.max_batch = max_batch

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.iwbc_batch_free(self._h)
            self._h = None

    def update(self, q, v, refs=None, contact_mask=None, out=None, fields=Results.FIELDS):
        """Host arrays in, host arrays out (the reference-facing call: H2D + kernels + D2H)."""
        P = self.problem
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, P.nq)
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, P.nv)
        B = q.shape[0]
        if refs is None:
            refs = P.default_refs()
        refs = np.ascontiguousarray(refs, dtype=np.float64)
        stride = 0 if refs.ndim == 1 else P.nref
        cm = None if contact_mask is None else np.ascontiguousarray(np.broadcast_to(np.asarray(contact_mask, dtype=np.uint32), (B,)))
        if out is None:
            out = Results(P, B, fields)
        st = out.as_struct()
        rc = lib().iwbc_batch_update(self._h, B, q.ctypes.data, v.ctypes.data, refs.ctypes.data, stride, _ptr(cm), C.byref(st))
        if rc != 0:
            raise IwbcError(_err(lib()))
        return out

    def update_device(self, B, q_ptr, v_ptr, refs_ptr, refs_stride, contact_mask, out_ptrs, stream=0):
        """All pointers are device addresses (ints); asynchronous on `stream`."""
        st = _Results(*[out_ptrs.get(f) for f in Results.FIELDS])
        rc = lib().iwbc_batch_update_device(self._h, B, q_ptr, v_ptr, refs_ptr, refs_stride, contact_mask, C.byref(st), stream)
        if rc != 0:
            raise IwbcError(_err(lib()))

    def rollout_device(self, B, q_ptr, v_ptr, refs_ptr, refs_stride, contact_mask, ticks, out_ptrs=None, stream=0):
        st = _Results(*[(out_ptrs or {}).get(f) for f in Results.FIELDS])
        rc = lib().iwbc_batch_rollout_device(self._h, B, q_ptr, v_ptr, refs_ptr, refs_stride, contact_mask, ticks, C.byref(st), stream)
        if rc != 0:
            raise IwbcError(_err(lib()))

    def behavior_rollout_device(self, behavior, B, q_ptr, v_ptr, refs_ptr, phase_ptr, tick0, ticks, init_refs=True, out_ptrs=None, mask_ptr=None,
                                stream=0):
        """Closed device-side roll-out of a compiled behaviour; all pointers are device addresses."""
        st = _Results(*[(out_ptrs or {}).get(f) for f in Results.FIELDS])
        rc = lib().iwbc_batch_behavior_rollout_device(self._h, behavior.handle, B, q_ptr, v_ptr, refs_ptr, phase_ptr, tick0, ticks,
                                                      1 if init_refs else 0, C.byref(st), mask_ptr, stream)
        if rc != 0:
            raise IwbcError(_err(lib()))

    def behavior_run(self, behavior, q, v, ticks, tick0=0, phase=None, out=None, fields=("tau", "status"), want_mask=False, trace=None):
        """`for tick: behavior.update()` with host arrays: q, v are advanced IN PLACE by `ticks` closed device-side ticks.
        trace: optional float64 array [ticks][B][nv] receiving every tick's command vector (controller->q(false))."""
        P = self.problem
        assert q.dtype == np.float64 and v.dtype == np.float64 and q.flags.c_contiguous and v.flags.c_contiguous
        B = q.reshape(-1, P.nq).shape[0]
        ph = None if phase is None else np.ascontiguousarray(phase, dtype=np.int32)
        if out is None:
            out = Results(P, B, fields)
        st = out.as_struct()
        mask = np.zeros(B, dtype=np.uint32) if want_mask else None
        if trace is None:
            rc = lib().iwbc_batch_behavior_run(self._h, behavior.handle, B, q.ctypes.data, v.ctypes.data, _ptr(ph), tick0, ticks, C.byref(st), _ptr(mask))
        else:
            assert trace.dtype == np.float64 and trace.flags.c_contiguous and trace.size == ticks * B * P.nv
            rc = lib().iwbc_batch_behavior_run_trace(self._h, behavior.handle, B, q.ctypes.data, v.ctypes.data, _ptr(ph), tick0, ticks, C.byref(st),
                                                     _ptr(mask), trace.ctypes.data)
        if rc != 0:
            raise IwbcError(_err(lib()))
        return (out, mask) if want_mask else out

    def check_collisions_device(self, collision_model, B, q_ptr, colliding_ptr, pair_ptr=None, stream=0):
        """Batched CollisionCheck::is_colliding on device-resident states."""
        rc = lib().iwbc_batch_check_collisions_device(self._h, collision_model.handle, B, q_ptr, colliding_ptr, pair_ptr, stream)
        if rc != 0:
            raise IwbcError(_err(lib()))

    def stabilizer(self):
        """Per-robot stabilizer state (filters, last normal forces, ...) for this batch: HumanoidPosTracker's sensor-driven pre-step."""
        return Stabilizer(self)

    def update_sensors(self, stab, q, v, sensors, refs=None, contact_mask=None, behavior_type=2, out=None, fields=Results.FIELDS, want_refs=False):
        """`HumanoidPosTracker::update(sensor_data)` for B robots, host arrays.  sensors[B][15] = lf_force rf_force lf_torque rf_torque
        imu_vel; behavior_type: int (everybody) or int array [B] (0 fixed base, 1 single support, 2 double support)."""
        P = self.problem
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, P.nq)
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, P.nv)
        B = q.shape[0]
        sensors = np.ascontiguousarray(sensors, dtype=np.float64).reshape(B, 15)
        refs = P.default_refs() if refs is None else np.ascontiguousarray(refs, dtype=np.float64)
        stride = 0 if refs.ndim == 1 else P.nref
        cm = None if contact_mask is None else np.ascontiguousarray(np.broadcast_to(np.asarray(contact_mask, dtype=np.uint32), (B,)))
        bt_all, bt = (int(behavior_type), None) if np.isscalar(behavior_type) else (2, np.ascontiguousarray(behavior_type, dtype=np.int32))
        if out is None:
            out = Results(P, B, fields)
        st = out.as_struct()
        used = np.zeros((B, P.nref)) if want_refs else None
        rc = lib().iwbc_batch_update_sensors(self._h, stab.handle, B, q.ctypes.data, v.ctypes.data, refs.ctypes.data, stride, _ptr(cm), _ptr(bt), bt_all,
                                             sensors.ctypes.data, C.byref(st), _ptr(used))
        if rc != 0:
            raise IwbcError(_err(lib()))
        return (out, used) if want_refs else out

    SAFETY_FIELDS = ("send_cmd", "colliding", "q_prev", "v_prev", "q_solver", "q_cmd", "tau_cmd", "dq_cmd", "ddq_cmd")

    def safe_rollout_device(self, collision_model, B, q_ptr, v_ptr, refs_ptr, refs_stride, contact_mask, ticks, safety_ptrs, status_ptr=None, stream=0):
        """Open-loop ticks with the controller's send_cmd latch / collision gating (controller.cpp:238-242, 263-282, 309-312); all device addresses."""
        s = _Safety(*[safety_ptrs.get(f) for f in self.SAFETY_FIELDS])
        rc = lib().iwbc_batch_safe_rollout_device(self._h, collision_model.handle if collision_model is not None else None, B, q_ptr, v_ptr, refs_ptr,
                                                  refs_stride, contact_mask, ticks, C.byref(s), status_ptr, stream)
        if rc != 0:
            raise IwbcError(_err(lib()))

    def qp_step_back_device(self, B, q_ptr, v_ptr, safety_ptrs, select_ptr=None, stream=0):
        s = _Safety(*[safety_ptrs.get(f) for f in self.SAFETY_FIELDS])
        rc = lib().iwbc_batch_qp_step_back_device(self._h, B, q_ptr, v_ptr, C.byref(s), select_ptr, stream)
        if rc != 0:
            raise IwbcError(_err(lib()))

    def debug_instance(self, q, v, ref=None, contact_mask=None):
        P = self.problem
        n = P.qp_shape(contact_mask)[0]
        ref = P.default_refs() if ref is None else np.ascontiguousarray(ref, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        M, nle, H, g, x = np.zeros((P.nv, P.nv)), np.zeros(P.nv), np.zeros((n, n)), np.zeros(n), np.zeros(n)
        m = P.default_contact_mask if contact_mask is None else contact_mask
        rc = lib().iwbc_batch_debug_instance(self._h, q.ctypes.data, v.ctypes.data, ref.ctypes.data, m, M.ctypes.data,
                                             nle.ctypes.data, H.ctypes.data, g.ctypes.data, x.ctypes.data)
        if rc != 0:
            raise IwbcError(_err(lib()))
        return dict(M=M, nle=nle, H=H, g=g, x=x)

    def objective_value(self, q, v, ref=None, contact_mask=None):
        """PosTracker::objective_value() of one robot: eiquadprog's objective x'Hx/2 + g'x at the solution of the QP at (q, v, ref)."""
        P = self.problem
        ref = P.default_refs() if ref is None else np.ascontiguousarray(ref, dtype=np.float64)
        q, v = np.ascontiguousarray(q, dtype=np.float64), np.ascontiguousarray(v, dtype=np.float64)
        m = P.default_contact_mask if contact_mask is None else contact_mask
        out = C.c_double(0.0)
        if lib().iwbc_batch_objective_value(self._h, q.ctypes.data, v.ctypes.data, ref.ctypes.data, m, C.byref(out)) != 0:
            raise IwbcError(_err(lib()))
        return out.value

    def task_cost(self, task, q, v, ddq, ref=None, contact_mask=None, tau=None):
        """Controller::cost(task) of one robot: norm of the task's own constraint residual at the accelerations ddq."""
        P = self.problem
        ref = P.default_refs() if ref is None else np.ascontiguousarray(ref, dtype=np.float64)
        q, v, ddq = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, v, ddq))
        tau = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
        m = P.default_contact_mask if contact_mask is None else contact_mask
        out = C.c_double(0.0)
        rc = lib().iwbc_batch_task_cost(self._h, q.ctypes.data, v.ctypes.data, ref.ctypes.data, m, task.encode(), ddq.ctypes.data,
                                        None if tau is None else tau.ctypes.data, C.byref(out))
        if rc != 0:
            raise IwbcError(_err(lib()))
        return out.value

    def debug_cycles(self):
        """Phase durations (SM cycles) of the last debug_instance call: dyn, asm, factor, equalities, active-set loop, decode."""
        t = np.zeros(7, dtype=np.int64)
        lib().iwbc_batch_debug_cycles(self._h, t.ctypes.data)
        return dict(zip(("dyn", "asm", "factor", "eq", "ineq", "decode"), np.diff(t).tolist()))

    def phase_profile(self, enable=True):
        """Read + clear the steady-state phase accumulators and switch profiling on/off. Returns mean SM cycles per
        instance of each phase since the last call, for the active-set kernel and for the fast kernel."""
        t = np.zeros(16, dtype=np.uint64)
        rc = lib().iwbc_batch_phase_profile(self._h, 1 if enable else 0, t.ctypes.data)
        if rc != 0:
            raise IwbcError(_err(lib()))
        out = {}
        if t[6]:
            out["active_set"] = dict(zip(("dyn", "asm", "factor", "eq", "ineq", "decode"), (t[:6] / float(t[6])).tolist()), instances=int(t[6]))
        if t[14]:
            out["fast"] = dict(zip(("dyn", "rows", "H_chol", "qr_solve", "check", "decode"), (t[8:14] / float(t[14])).tolist()), instances=int(t[14]))
        lp = np.zeros(16, dtype=np.uint64)
        lib().iwbc_batch_loop_profile(self._h, lp.ctypes.data)
        if lp[6]:
            out["loop"] = dict(zip(("pass", "load_n", "d", "z_r_coef", "step_rotate", "drops"), (lp[:6] / float(lp[6])).tolist()),
                               directions=int(lp[6]), adds=int(lp[7]), note="mean SM cycles per step direction",
                               z_phase=dict(zip(("r_ratio_last_warp", "coefficients_warp", "z_rows_warp0"), (lp[8:11] / float(lp[6])).tolist())))
        return out or None

    def fast_info(self):
        t = np.zeros(6, dtype=np.int64)
        lib().iwbc_batch_fast_info(self._h, t.ctypes.data)
        return dict(zip(("enabled", "fast_grid", "fast_smem", "slow_grid", "slow_smem", "threads"), t.tolist()))

    def launch_info(self):
        g, i, gr, sm = C.c_int(), C.c_int(), C.c_int(), C.c_long()
        lib().iwbc_batch_launch_info(self._h, C.byref(g), C.byref(i), C.byref(gr), C.byref(sm))
        return dict(threads_per_instance=g.value, instances_per_cta=i.value, grid=gr.value, smem_bytes=sm.value)

    @property
    def launch_count(self):
        return lib().iwbc_batch_launch_count(self._h)

    def enable_timing(self, on=True):
        lib().iwbc_batch_enable_timing(self._h, 1 if on else 0)

    def kernel_ms(self, reset=False):
        return lib().iwbc_batch_kernel_ms(self._h, 1 if reset else 0)


class MultiBatch:
    """One host call for a batch sharded over several GPUs of the node (contiguous shards, one worker thread per device)."""

    def __init__(self, problem, devices, max_batch_total):
        self.problem = problem
        devs = np.ascontiguousarray(devices, dtype=np.int32)
        self._h = lib().iwbc_multi_create(problem.handle, len(devs), devs.ctypes.data, max_batch_total)
        if not self._h:
            raise IwbcError(_err(lib()))
        self.ndev = len(devs)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.iwbc_multi_free(self._h)
            self._h = None

    def update(self, q, v, refs=None, contact_mask=None, out=None, fields=Results.FIELDS):
        P = self.problem
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, P.nq)
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, P.nv)
        B = q.shape[0]
        refs = P.default_refs() if refs is None else np.ascontiguousarray(refs, dtype=np.float64)
        stride = 0 if refs.ndim == 1 else P.nref
        cm = None if contact_mask is None else np.ascontiguousarray(np.broadcast_to(np.asarray(contact_mask, dtype=np.uint32), (B,)))
        if out is None:
            out = Results(P, B, fields)
        st = out.as_struct()
        rc = lib().iwbc_multi_update(self._h, B, q.ctypes.data, v.ctypes.data, refs.ctypes.data, stride, _ptr(cm), C.byref(st))
        if rc != 0:
            raise IwbcError(_err(lib()))
        return out

    @property
    def launch_count(self):
        return lib().iwbc_multi_launch_count(self._h)


def shard_range(total, rank, world):
    """Contiguous instance range [lo, hi) of shard `rank` of `world` (iwbc_shard_range: [g B / G, (g + 1) B / G))."""
    a, n = C.c_int(), C.c_int()
    if lib().iwbc_shard_range(total, world, rank, C.byref(a), C.byref(n)) != 0:
        raise IwbcError("bad shard")
    return a.value, a.value + n.value


def measure_fp64_tflops(device=0):
    r = lib().iwbc_measure_fp64_tflops(device)
    if r < 0:
        raise IwbcError(_err(lib()))
    return r


def qp_solve_batch(H, g, CE, ce0, CI, ci0, device=0):
    """Batched SolverHQPBase::solve (EiquadprogFast conventions). Arrays: H[B,n,n] g[B,n] CE[B,me,n] ce0[B,me] CI[B,mi,n] ci0[B,mi]."""
    H = np.ascontiguousarray(H, dtype=np.float64)
    B, n, _ = H.shape
    g = np.ascontiguousarray(g, dtype=np.float64)
    me = 0 if CE is None else CE.shape[1]
    mi = 0 if CI is None else CI.shape[1]
    CE = None if me == 0 else np.ascontiguousarray(CE, dtype=np.float64)
    ce0 = None if me == 0 else np.ascontiguousarray(ce0, dtype=np.float64)
    CI = None if mi == 0 else np.ascontiguousarray(CI, dtype=np.float64)
    ci0 = None if mi == 0 else np.ascontiguousarray(ci0, dtype=np.float64)
    x = np.zeros((B, n))
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    active = np.zeros((B, ACTIVE_WORDS), dtype=np.uint32)
    rc = lib().iwbc_qp_solve_batch(device, B, n, me, mi, H.ctypes.data, g.ctypes.data, _ptr(CE), _ptr(ce0), _ptr(CI), _ptr(ci0),
                                   x.ctypes.data, status.ctypes.data, iters.ctypes.data, active.ctypes.data)
    if rc != 0:
        raise IwbcError(_err(lib()))
    return x, status, iters, active
