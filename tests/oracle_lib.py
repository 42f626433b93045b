"""ctypes binding of the CPU oracle (oracle/liboracle.so) -- test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_ACTIVE_WORDS = 12
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        src = os.path.join(ORACLE_DIR, "wbc_oracle.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        _lib = C.CDLL(so)
        for f in ("oracle_cycle_batch", "oracle_rollout", "oracle_cycle_debug", "oracle_qp_solve", "oracle_rnea",
                  "oracle_frame", "oracle_problem_size"):
            getattr(_lib, f).restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def cycle_batch(problem, q, v, refs=None, contact_mask=None, nthreads=1):
    P = problem
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, P.nq)
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, P.nv)
    B = q.shape[0]
    refs = P.default_refs() if refs is None else np.ascontiguousarray(refs, dtype=np.float64)
    stride = 0 if refs.ndim == 1 else P.nref
    cm = None if contact_mask is None else np.ascontiguousarray(np.broadcast_to(np.asarray(contact_mask, dtype=np.uint32), (B,)))
    out = dict(tau=np.zeros((B, P.na)), dv=np.zeros((B, P.nv)), f=np.zeros((B, 12 * P.ncontacts)), q_next=np.zeros((B, P.nq)),
               v_next=np.zeros((B, P.nv)), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32),
               active=np.zeros((B, ORACLE_ACTIVE_WORDS), np.uint32), q_cmd=np.zeros((B, P.nv)), momentum=np.zeros((B, 3)))
    rc = lib().oracle_cycle_batch(C.c_void_p(P.data_ptr), C.c_int(B), _p(q), _p(v), _p(refs), C.c_long(stride), _p(cm),
                                  _p(out["tau"]), _p(out["dv"]), _p(out["f"]), _p(out["q_next"]), _p(out["v_next"]),
                                  _p(out["status"]), _p(out["iters"]), _p(out["active"]), _p(out["q_cmd"]), _p(out["momentum"]),
                                  C.c_int(nthreads))
    assert rc == 0
    return out


def rollout(problem, q, v, refs_t, ticks, contact_mask=None):
    P = problem
    q = np.array(q, dtype=np.float64)
    v = np.array(v, dtype=np.float64)
    refs_t = np.ascontiguousarray(refs_t, dtype=np.float64)
    stride = 0 if refs_t.ndim == 1 else P.nref
    m = P.default_contact_mask if contact_mask is None else contact_mask
    q_cmd = np.zeros((ticks, P.nv))
    tau = np.zeros((ticks, P.na))
    status = np.zeros(ticks, np.int32)
    rc = lib().oracle_rollout(C.c_void_p(P.data_ptr), C.c_int(ticks), _p(q), _p(v), _p(refs_t), C.c_long(stride), C.c_uint(m),
                              _p(q_cmd), _p(tau), _p(status))
    return dict(rc=rc, q=q, v=v, q_cmd=q_cmd, tau=tau, status=status)


def cycle_debug(problem, q, v, ref=None, contact_mask=None):
    P = problem
    m = P.default_contact_mask if contact_mask is None else contact_mask
    n, me, mi, _ = P.qp_shape(m)
    q = np.ascontiguousarray(q, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    ref = P.default_refs() if ref is None else np.ascontiguousarray(ref, dtype=np.float64)
    nj = 64
    o = dict(M=np.zeros((P.nv, P.nv)), nle=np.zeros(P.nv), J=np.zeros((6, P.nv)), Jcom=np.zeros((3, P.nv)), Ag=np.zeros((6, P.nv)),
             com=np.zeros(9), hg=np.zeros(12), H=np.zeros((n, n)), g=np.zeros(n), CE=np.zeros((me, n)), ce0=np.zeros(me),
             CI=np.zeros((mi, n)), ci0=np.zeros(mi), x=np.zeros(n), sizes=np.zeros(3, np.int32), status=np.zeros(1, np.int32),
             oMi=np.zeros((nj, 12)))
    rc = lib().oracle_cycle_debug(C.c_void_p(P.data_ptr), _p(q), _p(v), _p(ref), C.c_uint(m), _p(o["M"]), _p(o["nle"]), _p(o["J"]),
                                  _p(o["Jcom"]), _p(o["Ag"]), _p(o["com"]), _p(o["hg"]), _p(o["H"]), _p(o["g"]), _p(o["CE"]),
                                  _p(o["ce0"]), _p(o["CI"]), _p(o["ci0"]), _p(o["x"]), _p(o["sizes"]), _p(o["status"]), _p(o["oMi"]))
    assert rc == 0
    assert tuple(o["sizes"]) == (n, me, mi), (tuple(o["sizes"]), (n, me, mi))
    return o


def qp_solve(H, g, CE, ce0, CI, ci0):
    n = len(g)
    me = 0 if CE is None else len(ce0)
    mi = 0 if CI is None else len(ci0)
    H = np.ascontiguousarray(H, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    CE = np.zeros((0, n)) if me == 0 else np.ascontiguousarray(CE, dtype=np.float64)
    ce0 = np.zeros(0) if me == 0 else np.ascontiguousarray(ce0, dtype=np.float64)
    CI = np.zeros((0, n)) if mi == 0 else np.ascontiguousarray(CI, dtype=np.float64)
    ci0 = np.zeros(0) if mi == 0 else np.ascontiguousarray(ci0, dtype=np.float64)
    x = np.zeros(n)
    active = -np.ones(mi + 1, np.int32)
    iters = np.zeros(1, np.int32)
    lam = np.zeros(me + mi)
    st = lib().oracle_qp_solve(C.c_int(n), C.c_int(me), C.c_int(mi), _p(H), _p(g), _p(CE), _p(ce0), _p(CI), _p(ci0), _p(x),
                               _p(active), _p(iters), _p(lam))
    act = []
    for a in active[:mi]:
        if a < 0:
            break
        act.append(int(a))
    return dict(status=st, x=x, active=act, iters=int(iters[0]), lam=lam)


def rnea(problem, q, v, a):
    tau = np.zeros(problem.nv)
    lib().oracle_rnea(C.c_void_p(problem.data_ptr), _p(np.ascontiguousarray(q, dtype=np.float64)),
                      _p(np.ascontiguousarray(v, dtype=np.float64)), _p(np.ascontiguousarray(a, dtype=np.float64)), _p(tau))
    return tau


def frame(problem, q, v, frame_id):
    oMf, vel, acc, J = np.zeros(12), np.zeros(6), np.zeros(6), np.zeros((6, problem.nv))
    lib().oracle_frame(C.c_void_p(problem.data_ptr), _p(np.ascontiguousarray(q, dtype=np.float64)),
                       _p(np.ascontiguousarray(v, dtype=np.float64)), C.c_int(frame_id), _p(oMf), _p(vel), _p(acc), _p(J))
    return dict(oMf=oMf, vel=vel, acc=acc, J=J)


def stab_model(problem, q, v):
    """What the stabilizer reads from tsid's data: CoM sample (9), the two ankle (contact joint) positions, imu_link's angular velocity."""
    out = np.zeros(21)
    lib().oracle_stab_model(C.c_void_p(problem.data_ptr), _p(np.ascontiguousarray(q, dtype=np.float64)),
                            _p(np.ascontiguousarray(v, dtype=np.float64)), _p(out))
    return dict(com=out[0:9].copy(), lf_p=out[9:12].copy(), rf_p=out[12:15].copy(), imu_w=out[15:18].copy())


# ----------------------------------------------------------------- CPU simulation of the device source (test tooling)
_sim = None


def sim_lib():
    global _sim
    if _sim is None:
        so = os.path.join(ROOT, "tests", "libcpusim.so")
        srcs = [os.path.join(ROOT, "tests", "cpu_sim.cpp"), os.path.join(ROOT, "inria_wbc_b200", "csrc", "cycle_kernel.cuh"),
                os.path.join(ROOT, "inria_wbc_b200", "csrc", "fast_kernel.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unused", "-o", so, srcs[0]])
        _sim = C.CDLL(so)
    return _sim


def sim_cycle_batch(problem, q, v, refs=None, contact_mask=None):
    P = problem
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, P.nq)
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, P.nv)
    B = q.shape[0]
    refs = P.default_refs() if refs is None else np.ascontiguousarray(refs, dtype=np.float64)
    stride = 0 if refs.ndim == 1 else P.nref
    m = P.default_contact_mask if contact_mask is None else contact_mask
    out = dict(tau=np.zeros((B, P.na)), dv=np.zeros((B, P.nv)), f=np.zeros((B, 12 * P.ncontacts)), q_next=np.zeros((B, P.nq)),
               v_next=np.zeros((B, P.nv)), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32),
               active=np.zeros((B, ORACLE_ACTIVE_WORDS), np.uint32), q_cmd=np.zeros((B, P.nv)), momentum=np.zeros((B, 3)))
    rc = sim_lib().sim_cycle_batch(C.c_void_p(P.data_ptr), C.c_int(B), _p(q), _p(v), _p(refs), C.c_long(stride), C.c_uint(m),
                                   _p(out["tau"]), _p(out["dv"]), _p(out["f"]), _p(out["q_next"]), _p(out["v_next"]), _p(out["status"]),
                                   _p(out["iters"]), _p(out["active"]), _p(out["q_cmd"]), _p(out["momentum"]))
    assert rc == 0
    return out


def sim_fast_batch(problem, q, v, refs=None, contact_mask=None):
    """fast_kernel.cuh (equality-constrained first pass) chained with the active-set kernel, host-compiled."""
    P = problem
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, P.nq)
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, P.nv)
    B = q.shape[0]
    refs = P.default_refs() if refs is None else np.ascontiguousarray(refs, dtype=np.float64)
    stride = 0 if refs.ndim == 1 else P.nref
    m = P.default_contact_mask if contact_mask is None else contact_mask
    out = dict(tau=np.zeros((B, P.na)), dv=np.zeros((B, P.nv)), f=np.zeros((B, 12 * P.ncontacts)), q_next=np.zeros((B, P.nq)),
               v_next=np.zeros((B, P.nv)), status=np.zeros(B, np.int32), iters=np.zeros(B, np.int32),
               active=np.zeros((B, ORACLE_ACTIVE_WORDS), np.uint32), q_cmd=np.zeros((B, P.nv)), momentum=np.zeros((B, 3)),
               fast=np.zeros(B, np.int32))
    rc = sim_lib().sim_fast_batch(C.c_void_p(P.data_ptr), C.c_int(B), _p(q), _p(v), _p(refs), C.c_long(stride), C.c_uint(m),
                                  _p(out["tau"]), _p(out["dv"]), _p(out["f"]), _p(out["q_next"]), _p(out["v_next"]), _p(out["status"]),
                                  _p(out["iters"]), _p(out["active"]), _p(out["q_cmd"]), _p(out["momentum"]), _p(out["fast"]))
    assert rc == 0, "fast path not applicable to this problem"
    return out


def sim_qp_solve(H, g, CE, ce0, CI, ci0):
    n = len(g)
    me = 0 if CE is None else len(ce0)
    mi = 0 if CI is None else len(ci0)
    H = np.ascontiguousarray(H, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    CE = np.zeros((1, n)) if me == 0 else np.ascontiguousarray(CE, dtype=np.float64)
    ce0 = np.zeros(1) if me == 0 else np.ascontiguousarray(ce0, dtype=np.float64)
    CI = np.zeros((1, n)) if mi == 0 else np.ascontiguousarray(CI, dtype=np.float64)
    ci0 = np.zeros(1) if mi == 0 else np.ascontiguousarray(ci0, dtype=np.float64)
    x = np.zeros(n)
    status = np.zeros(1, np.int32)
    iters = np.zeros(1, np.int32)
    active = np.zeros(ORACLE_ACTIVE_WORDS, np.uint32)
    sim_lib().sim_qp_solve(C.c_int(n), C.c_int(me), C.c_int(mi), _p(H), _p(g), _p(CE), _p(ce0), _p(CI), _p(ci0), _p(x), _p(status),
                           _p(iters), _p(active))
    return dict(status=int(status[0]), x=x, iters=int(iters[0]), active=active)


def collision(problem, q, spheres_p_radius, spheres_joint_member, nmembers):
    """oracle_collision: (hit, member a, i, member b, j) for one configuration."""
    pr = np.ascontiguousarray(spheres_p_radius, dtype=np.float64)
    jm = np.ascontiguousarray(spheres_joint_member, dtype=np.int32)
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.zeros(5, dtype=np.int32)
    rc = lib().oracle_collision(C.c_void_p(problem.data_ptr), _p(q), C.c_int(pr.shape[0]), _p(pr), _p(jm), C.c_int(nmembers), _p(out))
    assert rc == 0
    return out
