"""problem_loader.py -- the oracle's OWN reader of the reference's configuration files.  TEST INFRASTRUCTURE ONLY.

Builds the flat `iwbc_problem_t` (include/iwbc_problem.h, the C header shared by the product library and oracle/wbc_oracle.cpp)
from the controller YAML, the task YAML, frames.yaml, the SRDF and the URDF with PyYAML and xml.etree -- WITHOUT the product's host
code (inria_wbc_b200/csrc/stack.hpp, model.hpp, yaml_lite.hpp, xml_lite.hpp are never read, linked or imported here).  Two uses:

  * tests/test_oracle_loader.py compares this struct field by field with the one the product's config compiler emits
    (`iwbc_problem_data()`), so a mistake in gains, masks, level assignment, bound tightening, friction rows or URDF inertia
    folding on either side shows up instead of cancelling out in the GPU-vs-oracle tests;
  * `bench.py --impl reference` drives the CPU oracle from this loader, so that arm never loads libiwbc_b200.so.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

What it follows (reference file:line; upstream pinocchio / urdfdom / TSID behaviour is restated from their published algorithms,
SURVEY.md Appendix A -- none of them is vendored in /root/reference):
  Controller ctor ........... src/controllers/controller.cpp:65-130 (urdf, floating base -> JointModelFreeFlyer "root_joint", dt)
  PosTracker ctor ........... src/controllers/pos_tracker.cpp:40-159 (frames :52-55,191-209; SRDF posture :57-77 with the
                              quaternion normalised; first computeProblemData at v = 0 :105-106; parse_tasks in file order
                              :161-189; joint_range_reduction :115-148)
  task factories ............ src/controllers/tasks.cpp:38-404 (Kd = 2 sqrt(Kp); masks; level 0 for bounds / actuation bounds /
                              contacts, level 1 otherwise; contact corner points :353-356; w_force_feet tasks.hpp:23)
  self-collision constants .. src/tsid/task-self-collision.cpp:163-164
  HumanoidPosTracker ........ src/controllers/humanoid_pos_tracker.cpp:50-79 (init_com: CoM reference between the feet, 1 cm check)
  TalosPosTracker ........... src/controllers/talos_pos_tracker.cpp:41-59 (hip-roll range around the start posture; the division
                              by 180 is done in float)
  pinocchio::urdf::buildModel: depth-first over the link tree, a link's child joints in urdfdom's order (its joints live in a
                              std::map, i.e. sorted by joint name); fixed joints fold the child link into the supporting joint
                              (Inertia::operator+=) and leave a FIXED_JOINT and a BODY frame; every movable joint gets a JOINT frame
  pinocchio::srdf::loadReferenceConfigurations: neutral configuration overwritten by the listed joints
  Contact6d ................. force generator T = [I; skew(p_i)], 4-sided friction pyramid + summed normal force row,
                              regularisation weights (1, 1, 1e-3, 2, 2, 2)
"""
import ctypes as C
import math
import os
import xml.etree.ElementTree as ET

import numpy as np
import yaml

# ------------------------------------------------------------------------------------------------ iwbc_problem.h mirror
MAXJ, MAXNV, MAXF, MAXT, MAXC, MAXAV, MAXL0, MAXL1, MAXWV = 56, 64, 48, 32, 4, 12, 16, 40, 4
J_FREEFLYER, J_REVOLUTE, J_PRISMATIC = 0, 1, 2
T_SE3, T_COM, T_POSTURE, T_MOMENTUM, T_SELFCOL, T_BOUNDS, T_ACTBOUNDS, T_TORQUE, T_COP = range(9)
B_BASEDYN, B_CONTACT_MOTION, B_CONTACT_FORCE, B_CONTACT_REG, B_TASK = range(5)
REF_SE3, REF_COM, REF_MOM, REF_CONTACT = 24, 9, 18, 32
BT_FIXED_BASE, BT_SINGLE_SUPPORT, BT_DOUBLE_SUPPORT = 0, 1, 2


class TaskS(C.Structure):
    _fields_ = [("type", C.c_int), ("level", C.c_int), ("nrows", C.c_int), ("frame", C.c_int), ("ref_off", C.c_int), ("nav", C.c_int),
                ("av_frame", C.c_int * MAXAV), ("mask", C.c_ubyte * MAXNV), ("weight", C.c_double), ("kp", C.c_double * 6),
                ("kd", C.c_double * 6), ("av_r0", C.c_double * MAXAV), ("radius", C.c_double), ("sc_k", C.c_double),
                ("sc_sp", C.c_double), ("sc_m", C.c_double)]


class ContactS(C.Structure):
    _fields_ = [("frame", C.c_int), ("ref_off", C.c_int), ("active_default", C.c_int), ("pad_", C.c_int), ("kp", C.c_double * 6),
                ("kd", C.c_double * 6), ("T", C.c_double * 72), ("B", C.c_double * 204), ("lb", C.c_double * 17), ("ub", C.c_double * 17),
                ("wreg", C.c_double * 6), ("fref", C.c_double * 6), ("w_forcereg", C.c_double)]


class BlockS(C.Structure):
    _fields_ = [("kind", C.c_int), ("index", C.c_int)]


class StabS(C.Structure):
    _fields_ = [("filter_size", C.c_int), ("use_zmp", C.c_int), ("use_momentum", C.c_int), ("zmp_w_variant", C.c_int),
                ("com_gains", C.c_double * 6), ("ankle_gains", C.c_double * 6), ("ffda_gains", C.c_double * 3), ("pad_", C.c_double),
                ("zmp_p", C.c_double * 6), ("zmp_d", C.c_double * 6), ("zmp_w", C.c_double * 6), ("momentum_p", C.c_double * 6),
                ("momentum_d", C.c_double * 6)]


class ProblemS(C.Structure):
    _fields_ = [("magic", C.c_uint), ("version", C.c_uint), ("nj", C.c_int), ("nq", C.c_int), ("nv", C.c_int), ("na", C.c_int),
                ("floating_base", C.c_int), ("nframes", C.c_int), ("ntasks", C.c_int), ("ncontacts", C.c_int), ("nl0", C.c_int),
                ("nl1", C.c_int), ("nref", C.c_int), ("sc_jacobian_world", C.c_int), ("dt", C.c_double), ("gravity", C.c_double * 3),
                ("hess_reg", C.c_double), ("total_mass", C.c_double),
                ("jtype", C.c_int * MAXJ), ("parent", C.c_int * MAXJ), ("idx_q", C.c_int * MAXJ), ("idx_v", C.c_int * MAXJ),
                ("depth", C.c_int * MAXJ), ("last_desc", C.c_int * MAXJ), ("jplace", (C.c_double * 12) * MAXJ),
                ("axis", (C.c_double * 3) * MAXJ), ("mass", C.c_double * MAXJ), ("com", (C.c_double * 3) * MAXJ),
                ("inertia", (C.c_double * 6) * MAXJ),
                ("q_lo", C.c_double * MAXNV), ("q_hi", C.c_double * MAXNV), ("v_max", C.c_double * MAXNV), ("tau_max", C.c_double * MAXNV),
                ("fparent", C.c_int * MAXF), ("fplace", (C.c_double * 12) * MAXF),
                ("tasks", TaskS * MAXT), ("contacts", ContactS * MAXC), ("l0", BlockS * MAXL0), ("l1", BlockS * MAXL1),
                ("q0", C.c_double * (MAXNV + 1)),
                ("nwreg", C.c_int), ("pad2_", C.c_int), ("wreg_variants", (C.c_double * 6) * MAXWV),
                ("stabilizer_activated", C.c_int), ("stab_present", C.c_int), ("stab_talos", C.c_int), ("stab_task_com", C.c_int),
                ("stab_task_lf", C.c_int), ("stab_task_rf", C.c_int), ("stab_task_torso", C.c_int), ("stab_task_momentum", C.c_int),
                ("stab_contact_lf", C.c_int), ("stab_contact_rf", C.c_int), ("stab_imu_frame", C.c_int), ("pad3_", C.c_int),
                ("stab", StabS * 3)]


class ConfigError(RuntimeError):
    pass


# ------------------------------------------------------------------------------------------------ small rigid-body algebra
def _rpy(r, p, y):
    """urdfdom Rotation::setFromRPY: R = Rz(yaw) Ry(pitch) Rx(roll)"""
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _skew(p):
    return np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])


class SE3:
    def __init__(self, R=None, p=None):
        self.R = np.eye(3) if R is None else np.asarray(R, dtype=float)
        self.p = np.zeros(3) if p is None else np.asarray(p, dtype=float)

    def __mul__(self, o):
        return SE3(self.R @ o.R, self.p + self.R @ o.p)

    def flat12(self):
        """jplace / fplace layout: R row-major (9) then p (3)"""
        return list(self.R.reshape(9)) + list(self.p)

    def sample24(self):
        """TrajectorySample(12, 6) of trajs::to_sample(SE3): p, R column-major; zero velocity and acceleration"""
        return list(self.p) + list(self.R.T.reshape(9)) + [0.0] * 12


class Inertia:
    """pinocchio::InertiaTpl: mass, lever (CoM), rotational inertia about the CoM"""

    def __init__(self, m=0.0, c=None, I=None):
        self.m, self.c, self.I = float(m), (np.zeros(3) if c is None else np.asarray(c, float)), (np.zeros((3, 3)) if I is None else np.asarray(I, float))

    def moved(self, X):
        """X.act(Y): the same body expressed in the frame X maps from"""
        return Inertia(self.m, X.R @ self.c + X.p, X.R @ self.I @ X.R.T)

    def __add__(self, o):
        # InertiaTpl::__plus__: (ma + mb, weighted lever, Ia + Ib - ma mb / (ma + mb) * skewSquare(AB))
        if self.m == 0.0 and not self.I.any():
            return o
        if o.m == 0.0 and not o.I.any():
            return self
        mab = self.m + o.m
        ab = self.c - o.c
        return Inertia(mab, (self.m * self.c + o.m * o.c) / mab, self.I + o.I - (self.m * o.m / mab) * (_skew(ab) @ _skew(ab)))


# ------------------------------------------------------------------------------------------------ URDF -> joint tree
class _Joint:
    def __init__(self, name, jtype, parent, placement, axis=(0, 0, 1), lim=(0, 0, 0, 0)):
        self.name, self.type, self.parent, self.placement = name, jtype, parent, placement
        self.axis = np.asarray(axis, float)
        self.q_lo, self.q_hi, self.v_max, self.tau_max = lim
        self.Y = Inertia()
        self.idx_q = self.idx_v = 0


class _Frame:
    def __init__(self, name, parent, placement, ftype):
        self.name, self.parent, self.placement, self.type = name, parent, placement, ftype


def _floats(s, n=3, default=0.0):
    v = [float(x) for x in (s or "").split()]
    return v + [default] * (n - len(v))


class RobotModel:
    """What pinocchio::urdf::buildModel leaves in pinocchio::Model, as far as the controller uses it."""

    def __init__(self, urdf_path, floating_base):
        root = ET.parse(urdf_path).getroot()
        if root.tag != "robot":
            raise ConfigError("urdf: expected <robot> in " + urdf_path)
        links, ujoints = {}, {}
        for e in root:
            if e.tag == "link":
                Y = Inertia()
                ine = e.find("inertial")
                if ine is not None:
                    o = ine.find("origin")
                    xyz = _floats(o.get("xyz") if o is not None else "")
                    rpy = _floats(o.get("rpy") if o is not None else "")
                    m = float(ine.find("mass").get("value")) if ine.find("mass") is not None else 0.0
                    it = ine.find("inertia")
                    g = (lambda k: float(it.get(k, "0"))) if it is not None else (lambda k: 0.0)
                    I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
                    R = _rpy(*rpy)
                    Y = Inertia(m, xyz, R @ I @ R.T)  # pinocchio::urdf convertFromUrdf
                links[e.get("name")] = Y
            elif e.tag == "joint":
                o = e.find("origin")
                xyz = _floats(o.get("xyz") if o is not None else "")
                rpy = _floats(o.get("rpy") if o is not None else "")
                ax = e.find("axis")
                axis = np.array(_floats(ax.get("xyz")) if ax is not None else [1.0, 0.0, 0.0])
                lim = e.find("limit")
                lims = (0.0, 0.0, 0.0, 0.0) if lim is None else tuple(float(lim.get(k, "0")) for k in ("lower", "upper", "velocity", "effort"))
                ujoints[e.get("name")] = dict(type=e.get("type", "fixed"), parent=e.find("parent").get("link"), child=e.find("child").get("link"),
                                              origin=SE3(_rpy(*rpy), xyz), axis=axis / np.linalg.norm(axis), lims=lims)
        children = {u["child"] for u in ujoints.values()}
        roots = [n for n in links if n not in children]
        if len(roots) != 1:
            raise ConfigError("urdf: expected exactly one root link, found %r" % roots)
        self.joints, self.frames = [], [_Frame("universe", -1, SE3(), "fixed")]
        self.nq = self.nv = 0
        self.floating_base = bool(floating_base)
        sup = -1
        if floating_base:  # buildModel(urdf, JointModelFreeFlyer(), model): joint "root_joint" under the universe
            j = _Joint("root_joint", J_FREEFLYER, -1, SE3())
            self.joints.append(j)
            self.nq, self.nv = 7, 6
            self.frames.append(_Frame("root_joint", 0, SE3(), "joint"))
            sup = 0
        by_parent = {}
        for name in sorted(ujoints):  # urdfdom keeps joints in a std::map: a link's child joints come sorted by joint name
            by_parent.setdefault(ujoints[name]["parent"], []).append(name)

        def visit(link, sup, placement):
            self.frames.append(_Frame(link, sup, placement, "body"))
            if sup >= 0:
                self.joints[sup].Y = self.joints[sup].Y + links[link].moved(placement)
            for jn in by_parent.get(link, []):
                u = ujoints[jn]
                here = placement * u["origin"]
                if u["type"] == "fixed":
                    self.frames.append(_Frame(jn, sup, here, "fixed"))
                    visit(u["child"], sup, here)
                elif u["type"] in ("revolute", "prismatic"):
                    j = _Joint(jn, J_PRISMATIC if u["type"] == "prismatic" else J_REVOLUTE, sup, here, u["axis"], u["lims"])
                    j.idx_q, j.idx_v = self.nq, self.nv
                    self.nq += 1
                    self.nv += 1
                    self.joints.append(j)
                    jid = len(self.joints) - 1
                    self.frames.append(_Frame(jn, jid, SE3(), "joint"))
                    visit(u["child"], jid, SE3())
                else:
                    raise ConfigError("urdf: joint type '%s' (%s) is outside the batched path" % (u["type"], jn))

        visit(roots[0], sup, SE3())
        self.reference_configurations = {}

    @property
    def na(self):
        return self.nv - 6 if self.floating_base else self.nv

    def joint_id(self, name):
        for i, j in enumerate(self.joints):
            if j.name == name:
                return i
        return -1

    def frame_id(self, name):
        for i, f in enumerate(self.frames):
            if f.name == name:
                return i
        return -1

    def exist_body(self, name):
        return any(f.name == name and f.type == "body" for f in self.frames)

    def actuated_names(self):
        return [j.name for j in self.joints if j.type != J_FREEFLYER]

    def load_srdf(self, path):
        for gs in ET.parse(path).getroot().iter("group_state"):
            q = np.zeros(self.nq)
            if self.floating_base:
                q[6] = 1.0  # pinocchio::neutral
            for j in gs.iter("joint"):
                jid = self.joint_id(j.get("name"))
                if jid < 0:
                    continue  # not in the model (e.g. the wheels with a no-wheel URDF)
                vals = [float(x) for x in j.get("value").split()]
                n = 7 if self.joints[jid].type == J_FREEFLYER else 1
                if len(vals) != n:
                    raise ConfigError("srdf: joint %s expects %d values" % (j.get("name"), n))
                q[self.joints[jid].idx_q:self.joints[jid].idx_q + n] = vals
            self.reference_configurations[gs.get("name")] = q

    # construction-time forward kinematics at v = 0 (pos_tracker.cpp:105-106)
    def forward_kinematics(self, q):
        oMi, mass, mc = [], 0.0, np.zeros(3)
        for j in self.joints:
            if j.type == J_FREEFLYER:
                x, y, z, w = q[j.idx_q + 3:j.idx_q + 7]
                R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                              [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                              [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
                M = SE3(R, q[j.idx_q:j.idx_q + 3])
            elif j.type == J_REVOLUTE:
                a, th = j.axis, q[j.idx_q]
                K = _skew(a)
                M = SE3(np.eye(3) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K))  # Rodrigues
            else:
                M = SE3(None, j.axis * q[j.idx_q])
            li = j.placement * M
            oMi.append(oMi[j.parent] * li if j.parent >= 0 else li)
            mc += j.Y.m * (oMi[-1].R @ j.Y.c + oMi[-1].p)
            mass += j.Y.m
        return oMi, mc / mass, mass

    def frame_placement(self, oMi, fid):
        f = self.frames[fid]
        return oMi[f.parent] * f.placement if f.parent >= 0 else f.placement


# ------------------------------------------------------------------------------------------------ YAML helpers (yaml-cpp `as<T>()`)
def _load_yaml(path):
    with open(path) as fh:
        return yaml.load(fh, Loader=yaml.BaseLoader)  # every scalar stays a string; conversions are explicit, as with yaml-cpp


def _as_bool(s):
    if str(s).lower() in ("true", "yes", "on", "y", "1"):
        return True
    if str(s).lower() in ("false", "no", "off", "n", "0"):
        return False
    raise ConfigError("not a boolean: %r" % s)


def _need(node, key, where):
    if not isinstance(node, dict) or key not in node:
        raise ConfigError("config error reading %s.%s" % (where, key))
    return node[key]


def _mask(s, n, what):
    s = str(s)
    if len(s) != n:
        raise ConfigError("%s mask needs %d entries, here: %s" % (what, n, s))
    return [0 if ch == "0" else 1 for ch in s]


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)  # (only CONTROLLER.stabilizer.activated-style overrides use this)
        else:
            dst[k] = v


def _stringify(v):
    if isinstance(v, dict):
        return {k: _stringify(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_stringify(x) for x in v]
    if isinstance(v, bool):
        return "true" if v else "false"
    return str(v)


# ------------------------------------------------------------------------------------------------ the problem
class OracleProblem:
    """Duck-types the read-only part of inria_wbc_b200.Problem that tests/oracle_lib.py and bench.py's CPU legs use."""

    def __init__(self, controller_yaml, robots_path, etc_path, overrides=None):
        self.S = ProblemS()
        cfg = _load_yaml(controller_yaml)
        c = _need(cfg, "CONTROLLER", os.path.basename(controller_yaml))
        if overrides:
            _merge(c, _stringify(overrides))
        self.controller_name = str(_need(c, "name", "CONTROLLER"))
        base_path = str(_need(c, "base_path", "CONTROLLER"))
        if base_path.startswith("@ETC@"):
            base_path = etc_path + base_path[5:]
        if str(_need(c, "solver", "CONTROLLER")) != "eiquadprog":
            raise ConfigError("solver in configuration file must be 'eiquadprog' on this path")
        S = self.S
        S.magic, S.version = 0x43425749, 2
        S.dt = float(_need(c, "dt", "CONTROLLER"))
        fb = _as_bool(_need(c, "floating_base", "CONTROLLER"))
        self.mimic_names = [str(x) for x in _need(c, "mimic_dof_names", "CONTROLLER")]
        self.closed_loop = _as_bool(_need(c, "closed_loop", "CONTROLLER"))
        urdf = str(_need(c, "urdf", "CONTROLLER"))
        model = self.model = RobotModel(urdf if urdf.startswith("/") else os.path.join(robots_path, urdf), fb)
        if "frames" in c:  # PosTracker::parse_frames
            for name, node in _load_yaml(os.path.join(base_path, str(c["frames"]))).items():
                ref = model.frame_id(str(_need(node, "ref", "frames." + name)))
                if ref < 0:
                    raise ConfigError("virtual frame %s: unknown reference frame" % name)
                pos = [float(x) for x in _need(node, "pos", "frames." + name)]
                f = model.frames[ref]
                model.frames.append(_Frame(name, f.parent, f.placement * SE3(None, pos), "fixed"))
        if len(model.joints) > MAXJ or model.nv > MAXNV:
            raise ConfigError("model too large for the flat problem description")
        model.load_srdf(os.path.join(base_path, str(_need(c, "configurations", "CONTROLLER"))))
        ref_config = str(_need(c, "ref_config", "CONTROLLER"))
        if ref_config not in model.reference_configurations:
            raise ConfigError("The following reference config is not in ref_map : " + ref_config)
        q0 = model.reference_configurations[ref_config].copy()
        if fb:
            q0[3:7] /= math.sqrt(float(q0[3:7] @ q0[3:7]))  # quat.normalized()
        nj, na, nv = len(model.joints), model.na, model.nv
        S.nj, S.nq, S.nv, S.na, S.floating_base = nj, model.nq, nv, na, int(fb)
        S.gravity[2] = -9.81
        S.hess_reg = 1e-8
        S.sc_jacobian_world = 1 if str(c.get("self_collision_jacobian", "")) == "world" else 0
        for j, J in enumerate(model.joints):
            S.jtype[j], S.parent[j], S.idx_q[j], S.idx_v[j] = J.type, J.parent, J.idx_q, J.idx_v
            S.depth[j] = S.depth[J.parent] + 1 if J.parent >= 0 else 0
            S.jplace[j][:] = J.placement.flat12()
            S.axis[j][:] = list(J.axis)
            S.mass[j] = J.Y.m
            S.com[j][:] = list(J.Y.c)
            I = J.Y.I
            S.inertia[j][:] = [I[0, 0], I[0, 1], I[0, 2], I[1, 1], I[1, 2], I[2, 2]]
            S.total_mass += J.Y.m
        for j in range(nj):  # joints are numbered depth first: a subtree is a contiguous index range
            last = j
            for k in range(j + 1, nj):
                a = k
                while a > j:
                    a = S.parent[a]
                if a == j:
                    last = k
            S.last_desc[j] = last
        act = [J for J in model.joints if J.type != J_FREEFLYER]
        for k, J in enumerate(act):
            S.q_lo[k], S.q_hi[k], S.v_max[k], S.tau_max[k] = J.q_lo, J.q_hi, J.v_max, J.tau_max
        S.q0[:model.nq] = list(q0)
        oMi, com, _ = model.forward_kinematics(q0)

        self.ref0, self.task_names, self.contact_names, self.frame_names = [], [], [], []

        def use_frame(name):
            if name in self.frame_names:
                return self.frame_names.index(name)
            fid = model.frame_id(name)
            if fid < 0:
                raise ConfigError("Unknown frame or joint [%s]" % name)
            k = len(self.frame_names)
            if k >= MAXF:
                raise ConfigError("too many task frames")
            self.frame_names.append(name)
            S.fparent[k] = model.frames[fid].parent
            S.fplace[k][:] = model.frames[fid].placement.flat12()
            S.nframes = k + 1
            return k

        def alloc(n, values=None):
            off = len(self.ref0)
            self.ref0 += [0.0] * n if values is None else list(values) + [0.0] * (n - len(values))
            return off

        if fb:  # InverseDynamicsFormulationAccForce: the base-dynamics equality opens level 0
            S.l0[S.nl0].kind, S.l0[S.nl0].index = B_BASEDYN, 0
            S.nl0 += 1
        task_file = str(_need(c, "tasks", "CONTROLLER"))
        tasks = _load_yaml(task_file if task_file.startswith("/") else os.path.join(base_path, task_file))
        bounds_task = -1

        def new_task(name, ttype, level):
            if S.ntasks >= MAXT:
                raise ConfigError("too many tasks")
            t = S.tasks[S.ntasks]
            t.type, t.level, t.ref_off, t.frame = ttype, level, -1, -1
            self.task_names.append(name)
            blk = S.l0[S.nl0] if level == 0 else S.l1[S.nl1]
            blk.kind, blk.index = B_TASK, S.ntasks
            if level == 0:
                S.nl0 += 1
            else:
                S.nl1 += 1
            S.ntasks += 1
            return t

        def gains(t, kp, n):
            for i in range(n):
                t.kp[i], t.kd[i] = kp, 2.0 * math.sqrt(kp)

        for name, node in tasks.items():  # parse_tasks: file order = HQP insertion order
            ttype = str(_need(node, "type", name))
            if ttype == "se3":
                kp, tracked = float(_need(node, "kp", name)), str(_need(node, "tracked", name))
                mask, weight = _mask(_need(node, "mask", name), 6, "SE3"), float(_need(node, "weight", name))
                joint, body = model.joint_id(tracked) >= 0, model.exist_body(tracked)
                if joint and body:
                    raise ConfigError("Ambiguous name to track for task %s: this is both a joint and a frame [%s]" % (name, tracked))
                if not joint and not body:
                    raise ConfigError("Unknown frame or joint [%s]" % tracked)
                t = new_task(name, T_SE3, 1)
                t.mask[:6], t.nrows, t.weight = mask, sum(mask), weight
                gains(t, kp, 6)
                t.frame = use_frame(tracked)
                ref = oMi[model.joint_id(tracked)] if joint else model.frame_placement(oMi, model.frame_id(tracked))
                t.ref_off = alloc(REF_SE3, ref.sample24())
            elif ttype == "com":
                kp, mask, weight = float(_need(node, "kp", name)), _mask(_need(node, "mask", name), 3, "CoM"), float(_need(node, "weight", name))
                t = new_task(name, T_COM, 1)
                t.mask[:3], t.nrows, t.weight = mask, sum(mask), weight
                gains(t, kp, 3)
                t.ref_off = alloc(REF_COM, com)
            elif ttype == "momentum":
                kp, weight, mask = float(_need(node, "kp", name)), float(_need(node, "weight", name)), _mask(_need(node, "mask", name), 6, "Momentum")
                t = new_task(name, T_MOMENTUM, 1)
                t.mask[:6], t.nrows, t.weight = mask, sum(mask), weight
                gains(t, kp, 6)
                t.ref_off = alloc(REF_MOM)
            elif ttype == "posture":
                kp, weight, ref_name = float(_need(node, "kp", name)), float(_need(node, "weight", name)), str(_need(node, "ref", name))
                if ref_name not in model.reference_configurations:
                    raise ConfigError("Reference name %s not found" % ref_name)
                t = new_task(name, T_POSTURE, 1)
                t.weight = weight
                t.kp[0], t.kd[0] = kp, 2.0 * math.sqrt(kp)
                mask = _mask(node["mask"], na, "posture") if "mask" in node else [1] * na
                t.mask[:na], t.nrows = mask, sum(mask)
                t.ref_off = alloc(3 * na, model.reference_configurations[ref_name][model.nq - na:])
            elif ttype == "bounds":
                t = new_task(name, T_BOUNDS, 0)
                t.weight, t.nrows = float(_need(node, "weight", name)), na
                bounds_task = S.ntasks - 1
            elif ttype == "cop":  # make_cop, tasks.cpp:156-178
                t = new_task(name, T_COP, 1)
                t.weight, t.nrows = float(_need(node, "weight", name)), 3
                t.ref_off = alloc(6, [0.0, 0.0, 0.0, 0.0, 0.0, 1.0])  # cop reference, contact normal
            elif ttype == "torque":  # make_torque, tasks.cpp:227-271
                t = new_task(name, T_TORQUE, 1)
                t.weight = float(_need(node, "weight", name))
                if "mask" in node:
                    if len(str(node["mask"])) != na:
                        raise ConfigError("wrong size in torque mask, expected:%d got:%d" % (na, len(str(node["mask"]))))
                    mask = _mask(node["mask"], na, "torque")
                else:
                    mask = [1] * na
                t.mask[:na], t.nrows = mask, sum(mask)
                scaling = [1.0] * na  # TaskActuationEquality's default weight vector
                if "scaling" in node:
                    scaling = [float(x) for x in node["scaling"]]
                    if len(scaling) != na:
                        raise ConfigError("wrong size in torque scaling, expected:%d got:%d" % (na, len(scaling)))
                t.ref_off = alloc(2 * na, [0.0] * na + scaling)  # tau_ref (zero), weight vector
            elif ttype == "actuation-bounds":
                t = new_task(name, T_ACTBOUNDS, 0)
                t.weight, t.nrows = float(_need(node, "weight", name)), na
            elif ttype == "self-collision":
                kp, kd, m = float(_need(node, "kp", name)), float(_need(node, "kd", name)), float(_need(node, "m", name))
                tracked, weight = str(_need(node, "tracked", name)), float(_need(node, "weight", name))
                margin, radius = float(_need(node, "margin", name)), float(_need(node, "radius", name))
                t = new_task(name, T_SELFCOL, 1)
                t.weight, t.nrows, t.radius, t.sc_m = weight, 1, radius, m
                t.kp[0], t.kd[0] = kp, kd
                t.sc_k = -math.log(math.pow(-1e-5 + 1., -1. / m) - 1.) / margin   # task-self-collision.cpp:163
                t.sc_sp = -1. / t.sc_k * math.log(-1 + math.pow(2, 1. / m))      # :164
                t.frame = use_frame(tracked)
                for av, r0 in _need(node, "avoided", name).items():  # (the reference walks an unordered_map: order fixed to the file's here)
                    if model.frame_id(av) < 0:
                        raise ConfigError("Frame %s in %s does not exists." % (av, name))
                    if t.nav >= MAXAV:
                        raise ConfigError("too many avoided frames in " + name)
                    t.av_frame[t.nav], t.av_r0[t.nav] = use_frame(av), float(r0)
                    t.nav += 1
            elif ttype == "contact":
                if S.ncontacts >= MAXC:
                    raise ConfigError("too many contacts")
                g = lambda k: float(_need(node, k, name))  # noqa: E731
                kp, jn = g("kp"), str(_need(node, "joint", name))
                lxn, lyn, lxp, lyp, lz, mu, fmin, fmax = (g(k) for k in ("lxn", "lyn", "lxp", "lyp", "lz", "mu", "fmin", "fmax"))
                normal = np.array([float(x) for x in _need(node, "normal", name)])
                if normal.size != 3:
                    raise ConfigError("normal size:%d" % normal.size)
                if model.frame_id(jn) < 0 or model.joint_id(jn) < 0:
                    raise ConfigError(jn + " does not exist!")
                cid = S.ncontacts
                ct = S.contacts[cid]
                S.ncontacts += 1
                self.contact_names.append(name)
                ct.frame, ct.active_default = use_frame(jn), 1
                gains(ct, kp, 6)
                pts = np.array([[-lxn, -lxn, lxp, lxp], [-lyn, lyp, -lyn, lyp], [lz, lz, lz, lz]])  # tasks.cpp:353-356
                T = np.zeros((6, 12))
                Bm = np.zeros((17, 12))
                # Contact6d: tangents t1 = n x ex (n x ey when degenerate), t2 = n x t1, normalised
                t1 = np.cross(normal, [1.0, 0.0, 0.0])
                if np.linalg.norm(t1) < 1e-5:
                    t1 = np.cross(normal, [0.0, 1.0, 0.0])
                t2 = np.cross(normal, t1)
                t1, t2 = t1 / np.linalg.norm(t1), t2 / np.linalg.norm(t2)
                for i in range(4):
                    T[0:3, 3 * i:3 * i + 3] = np.eye(3)
                    T[3:6, 3 * i:3 * i + 3] = _skew(pts[:, i])
                    Bm[4 * i + 0, 3 * i:3 * i + 3] = -t1 - mu * normal
                    Bm[4 * i + 1, 3 * i:3 * i + 3] = t1 - mu * normal
                    Bm[4 * i + 2, 3 * i:3 * i + 3] = -t2 - mu * normal
                    Bm[4 * i + 3, 3 * i:3 * i + 3] = t2 - mu * normal
                    Bm[16, 3 * i:3 * i + 3] = normal
                ct.T[:], ct.B[:] = list(T.reshape(-1)), list(Bm.reshape(-1))
                ct.lb[:], ct.ub[:] = [-1e10] * 16 + [fmin], [0.0] * 16 + [fmax]
                ct.wreg[:] = [1, 1, 1e-3, 2, 2, 2]
                ct.w_forcereg = 1e-3  # cst::w_force_feet
                ct.ref_off = alloc(REF_CONTACT, oMi[model.joint_id(jn)].sample24())  # then f_ref = 0 and weight vector 0 (Contact6d's default)
                for kind in (B_CONTACT_MOTION, B_CONTACT_FORCE):  # addRigidContact: motion + force at level 0, regularisation at level 1
                    S.l0[S.nl0].kind, S.l0[S.nl0].index = kind, cid
                    S.nl0 += 1
                S.l1[S.nl1].kind, S.l1[S.nl1].index = B_CONTACT_REG, cid
                S.nl1 += 1
            else:
                raise ConfigError("task type '%s' (%s) is not on the path the oracle restates" % (ttype, name))
        S.nref = len(self.ref0)
        S.nwreg = 1
        S.wreg_variants[0][:] = [1, 1, 1e-3, 2, 2, 2]

        if "joint_range_reduction" in c:  # pos_tracker.cpp:115-148
            red = float(c["joint_range_reduction"]) / 180 * math.pi
            if red < 0:
                raise ConfigError("Joint range reduction: reduction must be positive.")
            for i in range(na):
                if S.q_hi[i] - S.q_lo[i] < 2 * red:
                    raise ConfigError("Joint range reduction: reduction cannot be greater than actual range of motion")
                S.q_lo[i] += red
                S.q_hi[i] -= red
                if not (S.q_lo[i] <= q0[model.nq - na + i] <= S.q_hi[i]):
                    raise ConfigError("Joint range reduction: q0 falls outside reduced joint space.")
        humanoid = self.controller_name in ("humanoid-pos-tracker", "talos-pos-tracker")
        if humanoid:  # HumanoidPosTracker::init_com
            if "com" not in self.task_names:
                raise ConfigError("init_com: no com task")
            if not all(n in self.task_names for n in ("lf", "rf")) or not all(n in self.contact_names for n in ("contact_lfoot", "contact_rfoot")):
                raise ConfigError("init_com: contact_rfoot or contact_lfoot or lf or rf is missing, cannot compute initial com reference")
            co = S.tasks[self.task_names.index("com")].ref_off
            lo, ro = (S.tasks[self.task_names.index(n)].ref_off for n in ("lf", "rf"))
            final = [(self.ref0[lo] + self.ref0[ro]) / 2, (self.ref0[lo + 1] + self.ref0[ro + 1]) / 2]
            d = math.hypot(com[0] - final[0], com[1] - final[1])
            if d > 0.01:
                raise ConfigError("Wrong starting configuration: the CoM needs to be between the two feet with less than 1cm difference, but the distance is %g" % d)
            self.ref0[co:co + 2] = final
            # HumanoidPosTracker::parse_stabilizer (humanoid_pos_tracker.cpp:81-114) with parse_stab_conf (stabilizer_conf.hpp:46-92)
            sn = _need(c, "stabilizer", "CONTROLLER")
            S.stabilizer_activated = int(_as_bool(_need(sn, "activated", "CONTROLLER.stabilizer")))
            keys = ("params_fixed_base", "params_ss", "params_ds")  # behaviour types FIXED_BASE, SINGLE_SUPPORT, DOUBLE_SUPPORT
            if all(k in sn for k in keys):
                for bt, k in enumerate(keys):
                    st = _load_yaml(os.path.join(base_path, str(sn[k])))
                    sc = S.stab[bt]
                    for key, field, n in (("com", "com_gains", 6), ("ankle", "ankle_gains", 6), ("ffda", "ffda_gains", 3), ("zmp_p", "zmp_p", 6),
                                          ("zmp_d", "zmp_d", 6), ("zmp_w", "zmp_w", 6), ("momentum_p", "momentum_p", 6), ("momentum_d", "momentum_d", 6)):
                        vals = [float(x) for x in _need(st, key, str(sn[k]))]
                        if len(vals) != n:
                            raise ConfigError("you need %d coefficient in %s" % (n, key))
                        getattr(sc, field)[:] = vals
                    sc.use_zmp = int(_as_bool(_need(st, "use_zmp", str(sn[k]))))
                    sc.use_momentum = int(_as_bool(_need(st, "use_momentum", str(sn[k]))))
                    sc.filter_size = int(_need(st, "filter_size", str(sn[k])))
                    sc.zmp_w_variant = 0
                    if sc.use_zmp:  # setRegularizationTaskWeightVector(zmp_w): one more weight vector the contacts may carry
                        have = [list(S.wreg_variants[v]) for v in range(S.nwreg)]
                        w = list(sc.zmp_w)
                        if w not in have:
                            if S.nwreg >= MAXWV:
                                raise ConfigError("too many distinct zmp_w vectors")
                            S.wreg_variants[S.nwreg][:] = w
                            S.nwreg += 1
                            have.append(w)
                        sc.zmp_w_variant = have.index(w)
                S.stab_present = 1
            elif S.stabilizer_activated:
                raise ConfigError("config error reading CONTROLLER.stabilizer.params_*")
            tid = lambda n: self.task_names.index(n) if n in self.task_names else -1  # noqa: E731
            S.stab_task_com, S.stab_task_lf, S.stab_task_rf = tid("com"), tid("lf"), tid("rf")
            S.stab_task_torso, S.stab_task_momentum = tid("torso"), tid("momentum")
            S.stab_contact_lf, S.stab_contact_rf = self.contact_names.index("contact_lfoot"), self.contact_names.index("contact_rfoot")
            S.stab_talos = int(self.controller_name == "talos-pos-tracker")
            S.stab_imu_frame = use_frame("imu_link") if S.stab_talos and model.frame_id("imu_link") >= 0 else -1
        else:
            S.stab_task_com = S.stab_task_lf = S.stab_task_rf = S.stab_task_torso = S.stab_task_momentum = -1
            S.stab_contact_lf = S.stab_contact_rf = S.stab_imu_frame = -1
        if self.controller_name == "talos-pos-tracker":  # TalosPosTracker ctor
            roll = float(np.float32(_need(c, "torso_max_roll", "CONTROLLER")) / np.float32(180)) * math.pi
            names = model.actuated_names()
            for n in ("leg_left_2_joint", "leg_right_2_joint"):
                if n not in names:
                    raise ConfigError("Talos should have " + n)
                i = names.index(n)
                qi = q0[model.nq - na + i]
                if not (S.q_lo[i] <= qi <= S.q_hi[i]):
                    raise ConfigError("Error in bounds, the torso limits are not viable")
                S.q_lo[i], S.q_hi[i] = qi - roll, qi + roll
            if bounds_task < 0:
                raise ConfigError("talos-pos-tracker needs a 'bounds' task")
        self.ref0 = np.array(self.ref0)
        self.nq, self.nv, self.na, self.ncontacts, self.nref = S.nq, S.nv, S.na, S.ncontacts, S.nref
        self.default_contact_mask = sum(1 << k for k in range(S.ncontacts) if S.contacts[k].active_default)

    # ---- the read-only surface of inria_wbc_b200.Problem
    @property
    def data_ptr(self):
        return C.addressof(self.S)

    def struct(self):
        return self.S

    def default_refs(self):
        return self.ref0.copy()

    def q0(self):
        return np.array(self.S.q0[:self.nq])

    def names(self, kind):
        return {"dofs": self.model.actuated_names(), "tasks": list(self.task_names), "contacts": list(self.contact_names),
                "mimics": list(self.mimic_names)}[kind]

    def contact_index(self, name):
        return self.contact_names.index(name)

    def ref_slot(self, name):
        S = self.S
        if name in self.task_names:
            t = S.tasks[self.task_names.index(name)]
            if t.ref_off >= 0:
                return t.ref_off, {T_SE3: REF_SE3, T_COM: REF_COM, T_MOMENTUM: REF_MOM, T_TORQUE: 2 * S.na, T_COP: 6}.get(t.type, 3 * S.na)
        if name in self.contact_names:
            return S.contacts[self.contact_names.index(name)].ref_off, REF_CONTACT
        raise KeyError(name)

    def set_task_weight(self, name, w):
        self.S.tasks[self.task_names.index(name)].weight = float(w)

    def qp_shape(self, contact_mask=None):
        """nVar / nEq / nIn of InverseDynamicsFormulationAccForce and the number of level-1 rows"""
        S = self.S
        m = self.default_contact_mask if contact_mask is None else contact_mask
        on = [bool(m >> k & 1) for k in range(S.ncontacts)]
        n = S.nv + 12 * sum(on)
        me = mi = rows = 0
        for k in range(S.nl0):
            b = S.l0[k]
            if b.kind == B_BASEDYN:
                me += S.nv - S.na
            elif b.kind == B_CONTACT_MOTION:
                me += 6 if on[b.index] else 0
            elif b.kind == B_CONTACT_FORCE:
                mi += 34 if on[b.index] else 0
            else:
                mi += 2 * S.tasks[b.index].nrows
        for k in range(S.nl1):
            b = S.l1[k]
            rows += (6 if on[b.index] else 0) if b.kind == B_CONTACT_REG else S.tasks[b.index].nrows
        return n, me, mi, rows


def load(controller_yaml, overrides=None, robots_path=None, etc_path=None):
    """`controller_yaml` relative to the packaged etc/ (the reference's configuration files, imported unchanged)."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "inria_wbc_b200")
    etc_path = etc_path or os.path.join(root, "etc")
    robots_path = robots_path or os.path.join(root, "robots")
    if not os.path.isabs(controller_yaml) and not os.path.exists(controller_yaml):
        controller_yaml = os.path.join(etc_path, controller_yaml)
    return OracleProblem(controller_yaml, robots_path, etc_path, overrides)
