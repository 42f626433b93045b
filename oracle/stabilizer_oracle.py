"""stabilizer_oracle.py -- CPU restatement of the reference's sensor-driven pre-step.  TEST INFRASTRUCTURE ONLY.

Checker for the batched stabilizer; only tests/ may import it.  One `HumanoidController` = one robot of the
reference; every function follows the reference lines it cites, scalar code and all:

  Filter / MovingAverageFilter ............ include/inria_wbc/estimators/filtering.hpp:17-146
  Cop (per-foot and total CoP, filtered) ... src/estimators/cop.cpp:10-96, include/inria_wbc/estimators/cop.hpp:13-78
  com_to_zmp, com_admittance ............... src/stabilizers/stabilizer.cpp:9-59
  momentum_imu_admittance .................. src/stabilizers/stabilizer.cpp:107-140
  ankle_admittance ......................... src/stabilizers/stabilizer.cpp:168-224
  foot_force_difference_admittance ......... src/stabilizers/stabilizer.cpp:226-275
  zmp_distributor(_admittance) ............. src/stabilizers/stabilizer.cpp:277-311, 340-440
  HumanoidPosTracker::update ............... src/controllers/humanoid_pos_tracker.cpp:134-302 (set_behavior_type :116-132)
  TalosPosTracker::update (IMU momentum) ... src/controllers/talos_pos_tracker.cpp:125-151
  Eigen::Matrix3d::eulerAngles(0, 1, 2) .... upstream Eigen 3.3 (Geometry/EulerAngles.h), restated from its published algorithm

The control cycle itself, the model's CoM / ankle placements / IMU-frame velocity come from the C++ oracle (tests/oracle_lib.py).
The stabilizer rewrites references for ONE solve and puts them back afterwards (:291-301); here the behaviour's reference row is
simply left untouched and a modified copy goes into the solve.  What is NOT put back by the reference -- the contacts'
regularisation weight vector once the ZMP distributor has installed zmp_w (:220-221) -- persists here too.
"""
import math

import numpy as np

FMIN = 30.0  # estimators::FMIN (cop.hpp:13)
BT_FIXED_BASE, BT_SINGLE_SUPPORT, BT_DOUBLE_SUPPORT = 0, 1, 2


class MovingAverageFilter:
    """Filter + MovingAverageFilter (filtering.hpp:17-146)"""

    def __init__(self, nvar, wsize):
        self.nvar, self.wsize = nvar, wsize
        self.reset()

    def reset(self):  # :88-98
        self.buffer = np.zeros((self.nvar, self.wsize))
        self.filtered = np.zeros(self.nvar)
        self.cnt = 0

    def data_ready(self):  # :38
        return self.cnt >= self.wsize

    def filter(self, sample):  # :41-61
        sample = np.asarray(sample, float)
        if self.cnt < self.wsize:
            self.buffer[:, self.cnt] = sample
            self.cnt += 1
            self.filtered = self._mean(self.buffer[:, :self.cnt])
        else:
            if self.wsize > 1:
                self.buffer[:, :self.wsize - 1] = self.buffer[:, 1:self.wsize].copy()
            self.buffer[:, -1] = sample
            self.filtered = self._mean(self.buffer)
        return self.filtered.copy()

    @staticmethod
    def _mean(window):  # rowwise().mean(): sequential sum in column order, then the division
        acc = np.zeros(window.shape[0])
        for k in range(window.shape[1]):
            acc = acc + window[:, k]
        return acc / window.shape[1]

    def set_window_size(self, wsize):  # :63-86 (the count is NOT touched)
        if wsize != self.wsize:
            tmp = np.zeros((self.nvar, wsize))
            if wsize > self.wsize:
                tmp[:, :self.wsize] = self.buffer
            if wsize < self.wsize:
                tmp = self.buffer[:, self.wsize - wsize:].copy()
            self.buffer, self.wsize = tmp, wsize


class Cop:
    """estimators::Cop (cop.hpp, cop.cpp): None = boost::none"""

    def __init__(self, history_size=5):
        self.cop_filter, self.lcop_filter, self.rcop_filter = (MovingAverageFilter(2, history_size) for _ in range(3))
        self.cop_filtered = self.lcop_filtered = self.rcop_filtered = None

    def set_history_size(self, h):
        for f in (self.cop_filter, self.lcop_filter, self.rcop_filter):
            f.set_window_size(h)

    @staticmethod
    def foot_cop(foot_pos, torque, force):  # cop.cpp:86-95
        c = np.array(foot_pos[:2], float)
        c[0] += (-torque[1] - foot_pos[2] * force[0]) / force[2]
        c[1] += (torque[0] - foot_pos[2] * force[1]) / force[2]
        return c

    @staticmethod
    def total_cop(lf_pos, rf_pos, lcop, rcop, lf_force, rf_force):  # cop.cpp:69-82
        rl = lf_force[2] / (lf_force[2] + rf_force[2])
        rr = 1 - rl
        in_l = rl * lcop + rr * (rcop + rf_pos[:2] - lf_pos[:2])
        in_r = rl * (lcop + lf_pos[:2] - rf_pos[:2]) + rr * rcop
        return in_l * rl + in_r * rr

    def update(self, lf_pos, rf_pos, lf_torque, lf_force, rf_torque, rf_force, memory=True):  # cop.cpp:10-58
        if lf_force[2] > FMIN:
            lraw = self.foot_cop(lf_pos, lf_torque, lf_force)
            self.lcop_filtered = self.lcop_filter.filter(lraw)
        else:
            lraw = None
            self.lcop_filter.reset()
        if rf_force[2] > FMIN:
            rraw = self.foot_cop(rf_pos, rf_torque, rf_force)
            self.rcop_filtered = self.rcop_filter.filter(rraw)
        else:
            rraw = None
            self.rcop_filter.reset()
        if lraw is not None and rraw is not None:
            self.cop_filtered = self.cop_filter.filter(self.total_cop(lf_pos, rf_pos, lraw, rraw, lf_force, rf_force))
        elif memory:
            self.cop_filtered = None
        else:
            self.cop_filter.reset()
        if not self.cop_filter.data_ready():
            self.cop_filtered = None
        if not self.lcop_filter.data_ready():
            self.lcop_filtered = None
        if not self.rcop_filter.data_ready():
            self.rcop_filtered = None
        out = []
        for c in (self.cop_filtered, self.lcop_filtered, self.rcop_filtered):  # check_nan
            out.append(None if c is None or math.isnan(c[0]) or math.isnan(c[1]) else c)
        self.cop_filtered, self.lcop_filtered, self.rcop_filtered = out
        return out


# ---------------------------------------------------------------------------------------------- samples
def se3_of(sample24):
    """tsid::math::vectorToSE3 on TrajectorySample::getValue(): p(3), R column-major(9)"""
    return np.array(sample24[:3], float), np.array(sample24[3:12], float).reshape(3, 3).T


def pos12(p, R):
    return np.concatenate([p, R.T.reshape(9)])


def euler_angles_012(R):
    """Eigen::Matrix3d::eulerAngles(0, 1, 2): i, j, k = 0, 1, 2, `odd` = false; the first angle ends up in [0, pi]"""
    r0 = math.atan2(R[1, 2], R[2, 2])
    c2 = math.hypot(R[0, 0], R[0, 1])
    if r0 > 0.0:
        r0 -= math.pi
        r1 = math.atan2(-R[0, 2], -c2)
    else:
        r1 = math.atan2(-R[0, 2], c2)
    s1, c1 = math.sin(r0), math.cos(r0)
    r2 = math.atan2(s1 * R[2, 0] - c1 * R[1, 0], c1 * R[1, 1] - s1 * R[2, 1])
    return np.array([-r0, -r1, -r2])


def rot_xyz(e):
    """(AngleAxis(e0, X) * AngleAxis(e1, Y) * AngleAxis(e2, Z)).toRotationMatrix()"""
    c0, s0, c1, s1, c2, s2 = math.cos(e[0]), math.sin(e[0]), math.cos(e[1]), math.sin(e[1]), math.cos(e[2]), math.sin(e[2])
    Rx = np.array([[1, 0, 0], [0, c0, -s0], [0, s0, c0]])
    Ry = np.array([[c1, 0, s1], [0, 1, 0], [-s1, 0, c1]])
    Rz = np.array([[c2, -s2, 0], [s2, c2, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


# ---------------------------------------------------------------------------------------------- stabilizer.cpp
def com_to_zmp(com, acom):  # :9-20
    return com[:2] - com[2] / 9.81 * acom[:2]


def com_admittance(dt, p, cop, model_com9, com_ref9):  # :31-59; samples as value(3), derivative(3), second derivative(3)
    cor = com_to_zmp(model_com9[0:3], model_com9[6:9]) - cop
    out = np.array(com_ref9, float)
    out[0:2] -= p[0:2] * cor
    out[3:5] -= p[2:4] * cor / dt
    out[6:8] -= p[4:6] * cor / (dt * dt)
    return out


def momentum_imu_admittance(dt, p, d, imu_angular_vel, model_angular_vel, mom_ref18, max_ref=5.0):  # :107-140
    e = model_angular_vel - imu_angular_vel
    vec = np.array([e[0], e[1], e[2], e[1], e[0], e[2]])
    vel = np.clip(mom_ref18[6:12] - p * vec, -max_ref, max_ref)
    acc = np.clip(mom_ref18[12:18] - d * vec, -max_ref, max_ref)
    out = np.array(mom_ref18, float)
    out[6:12] = vel
    out[12:18] = acc / dt
    return out


def ankle_admittance(dt, p, cop_foot, model_foot_p, se3_ref24, contact_ref24):  # :168-224
    ap, aR = se3_of(se3_ref24)
    pitch = p[0] * (cop_foot[0] - model_foot_p[0])
    roll = p[1] * (cop_foot[1] - model_foot_p[1])
    e = euler_angles_012(aR)
    e[0] += roll
    e[1] += pitch
    Rn = rot_xyz(e)
    vel = np.array(se3_ref24[12:18], float)
    vel[4] += p[2] * pitch / dt
    vel[3] += p[3] * roll / dt
    acc = np.array(se3_ref24[18:24], float)
    acc[4] += p[4] * pitch / (dt * dt)
    acc[3] += p[5] * roll / (dt * dt)
    cp, _ = se3_of(contact_ref24)
    return np.concatenate([pos12(ap, Rn), vel, acc]), np.concatenate([pos12(cp, Rn), vel, acc])


def foot_force_difference_admittance(dt, Mg, p, lf_normal, rf_normal, lf_sensor, rf_sensor, torso_ref24):  # :226-275
    zctrl = abs(lf_normal - rf_normal) - abs(lf_sensor[2] - rf_sensor[2])
    torso_roll = p[0] * zctrl
    tp, tR = se3_of(torso_ref24)
    e = euler_angles_012(tR)
    e[0] += torso_roll
    vel = np.array(torso_ref24[12:18], float)
    vel[3] += p[1] * torso_roll / dt
    acc = np.array(torso_ref24[18:24], float)
    acc[3] += p[2] * torso_roll / (dt * dt)
    return np.concatenate([pos12(tp, rot_xyz(e)), vel, acc])


def closest_point_on_line(point, a, b):  # :431-440
    u = (b - a) / np.linalg.norm(b - a)
    return b + u * float((point - b) @ u)


def zmp_distributor(mass, zmp, lf_ref_p, rf_ref_p, lf_active, rf_active):  # :340-429
    PR, PL = np.zeros(2), np.zeros(2)
    if lf_active and rf_active:
        zmp3 = np.array([zmp[0], zmp[1], 0.0])
        first, second = np.array(lf_ref_p, float), np.array(rf_ref_p, float)
        PR, PL = second[:2].copy(), first[:2].copy()
        first[2] = second[2] = 0.0
        proj = closest_point_on_line(zmp3, first, second)
        n = np.linalg.norm
        if abs(n(second - first) - n(proj - first) - n(second - proj)) < 1e-10:
            alpha = n(proj - first) / n(first - second)
        elif n(proj - first) < n(proj - second):
            alpha = 0.0
        elif n(proj - first) > n(proj - second):
            alpha = 1.0
        else:
            raise RuntimeError("zmp_distributor this case should not exists")
    elif rf_active:
        alpha, PR = 1.0, np.array(rf_ref_p[:2], float)
    elif lf_active:
        alpha, PL = 0.0, np.array(lf_ref_p[:2], float)
    else:
        raise RuntimeError("No contact_rfoot or contact_rfoot in activated_contacts_ is talos jumping ?")
    f_right = -alpha * mass * 9.81
    f_left = -(1 - alpha) * mass * 9.81
    tau0 = -(PR - zmp) * f_right - (PL - zmp) * f_left
    tauR, tauL = np.zeros(2), np.zeros(2)
    tauR[1], tauL[1] = alpha * tau0[1], (1 - alpha) * tau0[1]
    if tau0[0] < 0:
        tauR[0], tauL[0] = tau0[0], 0.0
    else:
        tauR[0], tauL[0] = 0.0, tau0[0]
    left, right = np.zeros(6), np.zeros(6)
    left[2], left[3], left[4] = f_left, tauL[1], -tauL[0]
    right[2], right[3], right[4] = f_right, tauR[1], -tauR[0]
    return alpha, left, right


def zmp_distributor_admittance(dt, p, d, mass, lf_ref_p, rf_ref_p, lf_active, rf_active, zmp, model_com9):  # :277-311
    zmp_ref = com_to_zmp(model_com9[0:3], model_com9[6:9])
    _, l_tsid, r_tsid = zmp_distributor(mass, zmp_ref, lf_ref_p, rf_ref_p, lf_active, rf_active)
    _, l_sens, r_sens = zmp_distributor(mass, zmp, lf_ref_p, rf_ref_p, lf_active, rf_active)
    left = l_tsid - p * (l_tsid - l_sens) - d * (l_tsid - l_sens) / dt
    right = r_tsid - p * (r_tsid - r_sens) - d * (r_tsid - r_sens) / dt
    return left, right


# ---------------------------------------------------------------------------------------------- the controller
class HumanoidController:
    """HumanoidPosTracker (+ TalosPosTracker's IMU momentum admittance) for ONE robot.  `ol`: tests/oracle_lib; `problem`: anything
    with the read-only Problem surface (struct(), ref_slot(), nq / nv ...)."""

    def __init__(self, ol, problem, q=None, v=None):
        self.ol, self.P = ol, problem
        S = self.S = problem.struct()
        assert S.stab_present, "no stabilizer parameter sets in this problem"
        self.dt = S.dt
        self.q_tsid = problem.q0() if q is None else np.array(q, float)
        self.v_tsid = np.zeros(problem.nv) if v is None else np.array(v, float)
        self.data_q, self.data_v = problem.q0(), np.zeros(problem.nv)  # tsid's data: the constructor's computeProblemData (pos_tracker.cpp:105-106)
        self.behavior_type = BT_FIXED_BASE                              # humanoid_pos_tracker.cpp:39
        h = S.stab[self.behavior_type].filter_size                      # parse_stabilizer :100-113
        self.cop = Cop()
        self.cop.set_history_size(h)
        self.lf_force_filter, self.rf_force_filter, self.lf_torque_filter, self.rf_torque_filter, self.imu_filter = (
            MovingAverageFilter(3, h) for _ in range(5))
        self.lf_force_filtered, self.rf_force_filtered = np.zeros(3), np.zeros(3)
        self.contact_forces = {}      # activated_contacts_forces_ of the last solve: contact index -> f(12)
        self.wreg_variant = 0         # the contacts' regularisation weight vector (sticky once zmp_w is installed)
        self.base_fref = {S.stab_contact_lf: np.zeros(6), S.stab_contact_rf: np.zeros(6)}  # Contact6d::getForceReference() per contact
        self.off = {k: S.tasks[getattr(S, "stab_task_" + k)].ref_off if getattr(S, "stab_task_" + k) >= 0 else -1
                    for k in ("com", "lf", "rf", "torso", "momentum")}
        self.cl, self.cr = S.stab_contact_lf, S.stab_contact_rf
        self.coff = {self.cl: S.contacts[self.cl].ref_off, self.cr: S.contacts[self.cr].ref_off}

    def set_behavior_type(self, bt):  # humanoid_pos_tracker.cpp:116-132
        self.behavior_type = bt
        h = self.S.stab[bt].filter_size
        self.cop.set_history_size(h)
        for f in (self.lf_force_filter, self.rf_force_filter, self.lf_torque_filter, self.rf_torque_filter, self.imu_filter):
            f.set_window_size(h)

    def stabilized_refs(self, refs_row, sensors, contact_mask):
        """Everything of HumanoidPosTracker::update / TalosPosTracker::update that happens BEFORE _solve: returns the reference row
        the solve sees.  sensors: dict lf_force, rf_force, lf_torque, rf_torque, imu_vel (3 each)."""
        S, dt = self.S, self.dt
        conf = S.stab[self.behavior_type]
        r = np.array(refs_row, float)
        lf_active, rf_active = bool(contact_mask >> self.cl & 1), bool(contact_mask >> self.cr & 1)
        model = self.ol.stab_model(self.P, self.data_q, self.data_v)  # com(9), left / right ankle position, imu angular velocity
        if S.stab_talos and S.stabilizer_activated and conf.use_momentum:  # talos_pos_tracker.cpp:135-146
            imu_filtered = self.imu_filter.filter(sensors["imu_vel"])
            o = self.off["momentum"]
            # (the call passes the MODEL's velocity as `imu_angular_vel` and the filtered IMU as `model_angular_vel`: :143)
            r[o:o + 18] = momentum_imu_admittance(dt, np.array(conf.momentum_p), np.array(conf.momentum_d), model["imu_w"], imu_filtered, refs_row[o:o + 18])
        # humanoid_pos_tracker.cpp:161-177: the CoP estimate (all four F/T signals are present in the batched call)
        cops = self.cop.update(model["lf_p"], model["rf_p"], sensors["lf_torque"], sensors["lf_force"], sensors["rf_torque"], sensors["rf_force"])
        if not S.stabilizer_activated:
            return r
        # :187-203 force filters, only while the foot is loaded
        if np.linalg.norm(sensors["lf_force"]) > FMIN:
            self.lf_force_filtered = self.lf_force_filter.filter(sensors["lf_force"])
            self.lf_torque_filter.filter(sensors["lf_torque"])
        else:
            self.lf_force_filtered = np.zeros(3)
        if np.linalg.norm(sensors["rf_force"]) > FMIN:
            self.rf_force_filtered = self.rf_force_filter.filter(sensors["rf_force"])
            self.rf_torque_filter.filter(sensors["rf_torque"])
        else:
            self.rf_force_filtered = np.zeros(3)
        valid = cops[0] if cops[0] is not None else (cops[1] if cops[1] is not None else cops[2])  # :209
        # the contacts' force references as the contact objects hold them (:150-152; behaviours have no setter for them)
        for c in (self.cl, self.cr):
            r[self.coff[c] + 24:self.coff[c] + 30] = self.base_fref[c]
        oc = self.off["com"]
        if valid is not None:  # :211-214
            r[oc:oc + 9] = com_admittance(dt, np.array(conf.com_gains), valid, model["com"], refs_row[oc:oc + 9])
        if valid is not None and conf.use_zmp:  # :217-226
            lp, _ = se3_of(refs_row[self.coff[self.cl]:self.coff[self.cl] + 12])
            rp, _ = se3_of(refs_row[self.coff[self.cr]:self.coff[self.cr] + 12])
            left, right = zmp_distributor_admittance(dt, np.array(conf.zmp_p), np.array(conf.zmp_d), S.total_mass, lp, rp, lf_active, rf_active,
                                                     valid, model["com"])
            self.wreg_variant = conf.zmp_w_variant
            r[self.coff[self.cl] + 24:self.coff[self.cl] + 30] = left
            r[self.coff[self.cr] + 24:self.coff[self.cr] + 30] = right
        for c, active in ((self.cl, lf_active), (self.cr, rf_active)):
            r[self.coff[c] + 30] = float(self.wreg_variant)  # the installed weight vector is never put back (:291-301 restore references only)
            if not active:  # ... and the force reference only for the ACTIVATED contacts (:297-300): a lifted foot keeps what the distributor wrote
                self.base_fref[c] = r[self.coff[c] + 24:self.coff[c] + 30].copy()
        for cop_k, active, task, c, foot_p in ((cops[1], lf_active, "lf", self.cl, model["lf_p"]), (cops[2], rf_active, "rf", self.cr, model["rf_p"])):
            if cop_k is not None and active:  # :229-244
                o, oc2 = self.off[task], self.coff[c]
                se3_s, contact_s = ankle_admittance(dt, np.array(conf.ankle_gains), cop_k, foot_p, refs_row[o:o + 24], refs_row[oc2:oc2 + 24])
                r[o:o + 24] = se3_s
                r[oc2:oc2 + 24] = contact_s
        if self.cl in self.contact_forces and self.cr in self.contact_forces and self.lf_force_filter.data_ready() and self.rf_force_filter.data_ready():
            # :247-261 (normal forces of the PREVIOUS solve; Contact6d::getNormalForce = sum over the corners of n . f_i, n = contact normal)
            lf_n = self._normal_force(self.cl)
            rf_n = self._normal_force(self.cr)
            o = self.off["torso"]
            r[o:o + 24] = foot_force_difference_admittance(dt, S.total_mass * 9.81, np.array(conf.ffda_gains), lf_n, rf_n, self.lf_force_filtered,
                                                           self.rf_force_filtered, refs_row[o:o + 24])
        return r

    def _normal_force(self, c):
        """Contact6d::getNormalForce: sum over the corners of n . f_i (row 16 of the contact's force block is [n n n n])"""
        nrow = self.S.contacts[c].B[16 * 12:17 * 12]
        fn = 0.0
        for k in range(12):
            fn += nrow[k] * self.contact_forces[c][k]
        return fn

    def update(self, refs_row, sensors, contact_mask, q=None, v=None):
        """One HumanoidPosTracker::update: closed loop when (q, v) are given (:264-285), else the controller's own state."""
        q = self.q_tsid if q is None else np.array(q, float)
        v = self.v_tsid if v is None else np.array(v, float)
        r = self.stabilized_refs(refs_row, sensors, contact_mask)
        o = self.ol.cycle_batch(self.P, q[None], v[None], r, contact_mask=contact_mask)
        self.data_q, self.data_v = q.copy(), v.copy()  # computeProblemData ran on (q, v) whatever the solver says next
        if o["status"][0] != 0:
            raise RuntimeError("Controller failed, can't solve problem. Status : %d" % o["status"][0])
        self.q_tsid, self.v_tsid = o["q_next"][0], o["v_next"][0]
        self.contact_forces = {c: o["f"][0][12 * c:12 * c + 12].copy() for c in range(self.P.ncontacts) if contact_mask >> c & 1}
        return o, r
