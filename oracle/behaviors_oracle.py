"""behaviors_oracle.py -- numpy restatement of the reference's reference generators.  TEST INFRASTRUCTURE ONLY.

Checker for the device-side behaviours; only tests/, __graft_entry__.smoke() and bench.py's CPU legs
may import it.  Each function follows the reference lines it cites, scalar loops and all:

  minimum_jerk_polynom / min_jerk_trajectory / constant_traj .. include/inria_wbc/trajs/trajectory_generator.hpp:29-160
  humanoid::walk (state machine, contact add / remove) ......... src/behaviors/humanoid/walk.cpp:42-246

`Walk.tick(t)` returns what the reference's behaviour has written into the controller *before* the controller update of
global tick t: the full reference vector and the contact mask.
"""
import math

import numpy as np
import yaml

# States (include/inria_wbc/behaviors/humanoid/walk.hpp)
INIT, LF_INIT, LIFT_DOWN_LF, MOVE_COM_LEFT, LIFT_UP_RF, LIFT_DOWN_RF, MOVE_COM_RIGHT, LIFT_UP_LF, LIFT_DOWN_LF_FINAL, MOVE_COM_CENTER_FINAL = range(10)


def n_steps(dt, duration):
    return int(math.floor(duration / dt))  # trajectory_generator.hpp:73


def minimum_jerk_polynom(x0, xf, t, duration):
    """trajectory_generator.hpp:31-39 (order ZERO)"""
    td = t / duration
    td3 = math.pow(td, 3)
    return x0 + (xf - x0) * (6 * td * td * td3 - 15 * td * td3 + 10 * td3)


def min_jerk_vec(x0, xf, dt, duration):
    return [minimum_jerk_polynom(np.asarray(x0, float), np.asarray(xf, float), dt * i, duration) for i in range(n_steps(dt, duration))]


def constant_traj(p, dt, duration):
    return [p for _ in range(n_steps(dt, duration))]


class Walk:
    """src/behaviors/humanoid/walk.cpp.  `problem` is the compiled controller (for the construction-time references and the
    reference-vector layout); translations only change along the walk, rotations stay at their initial value (the waypoints
    are translates of the initial placements, walk.cpp:63-76)."""

    CHANNELS = ("com", "lf", "rf", "lh", "rh")

    def __init__(self, problem, config, num_of_cycles=None):
        c = config if isinstance(config, dict) else yaml.safe_load(open(config))
        c = c["BEHAVIOR"]
        self.P = problem
        self.dt = problem.struct().dt
        com_dur = float(np.float32(c["traj_com_duration"]))
        foot_dur = float(np.float32(c["traj_foot_duration"]))
        step_height = float(np.float32(c["step_height"]))
        step_length = float(np.float32(c["step_length"]))
        ncyc = int(c["num_of_cycles"]) if num_of_cycles is None else num_of_cycles
        if ncyc <= 0:
            raise ValueError("num_of_cycles needs to be more than 0")
        self.ref0 = problem.default_refs()
        self.slot = {k: problem.ref_slot(k)[0] for k in self.CHANNELS}
        self.cslot = {"lf": problem.ref_slot("contact_lfoot")[0], "rf": problem.ref_slot("contact_rfoot")[0]}
        self.bit = {"lf": 1 << problem.contact_index("contact_lfoot"), "rf": 1 << problem.contact_index("contact_rfoot")}
        dt = self.dt

        cycle = [INIT, LF_INIT]
        for _ in range(ncyc):
            cycle += [LIFT_DOWN_LF, MOVE_COM_LEFT, LIFT_UP_RF, LIFT_DOWN_RF, MOVE_COM_RIGHT, LIFT_UP_LF]
        cycle += [LIFT_DOWN_LF_FINAL, MOVE_COM_CENTER_FINAL]
        self.cycle = cycle

        def pos(name):
            o = self.slot[name]
            return self.ref0[o:o + 3].copy()

        def tr(p, v, idx):
            p2 = p.copy()
            p2[idx] += v
            return p2

        lf_low, rf_low = pos("lf"), pos("rf")
        lf_high, rf_high = tr(lf_low, step_height, 2), tr(rf_low, step_height, 2)
        com_init = pos("com")
        com_lf, com_rf = lf_low.copy(), rf_low.copy()
        com_lf[2] = com_init[2]
        com_rf[2] = com_init[2]
        lh_init, rh_init = pos("lh"), pos("rh")
        lh_forward, rh_forward = pos("lh"), pos("rh")
        T = {k: [] for k in self.CHANNELS}
        mj, ct = min_jerk_vec, constant_traj
        for s in cycle:
            if s == INIT:
                T["rf"].append(ct(rf_low, dt, com_dur)), T["lf"].append(ct(lf_low, dt, com_dur))
                T["com"].append(mj(com_init, com_rf, dt, com_dur))
                T["lh"].append(ct(lh_init, dt, com_dur)), T["rh"].append(ct(rh_init, dt, com_dur))
            elif s == LF_INIT:
                T["rf"].append(ct(rf_low, dt, foot_dur)), T["lf"].append(mj(lf_low, lf_high, dt, foot_dur))
                T["com"].append(ct(com_rf.copy(), dt, foot_dur))
                T["lh"].append(ct(lh_init, dt, foot_dur)), T["rh"].append(ct(rh_init, dt, foot_dur))
            elif s == LIFT_DOWN_LF:
                diff = np.float32(rf_low[0] - lf_low[0])  # `float diff`, float step_length_: a float sum (walk.cpp:97,116-118)
                adv = float(np.float32(diff + np.float32(step_length)))
                lf_low = tr(lf_low, adv, 0)
                lh_forward = tr(lh_init, adv, 0)
                T["rf"].append(ct(rf_low, dt, foot_dur)), T["lf"].append(mj(lf_high, lf_low, dt, foot_dur))
                T["com"].append(ct(com_rf.copy(), dt, foot_dur))
                T["lh"].append(mj(lh_init, lh_forward, dt, com_dur)), T["rh"].append(ct(rh_init, dt, com_dur))
                lh_init = lh_forward
            elif s == MOVE_COM_LEFT:
                com_lf = com_lf.copy()
                com_lf[0] = lf_low[0]
                T["rf"].append(ct(rf_low, dt, com_dur)), T["lf"].append(ct(lf_low, dt, com_dur))
                T["com"].append(mj(com_rf, com_lf, dt, com_dur))
                T["lh"].append(ct(lh_init, dt, com_dur)), T["rh"].append(ct(rh_init, dt, com_dur))
            elif s == LIFT_UP_RF:
                rf_high = rf_high.copy()
                rf_high[0] = lf_low[0]
                T["rf"].append(mj(rf_low, rf_high, dt, foot_dur)), T["lf"].append(ct(lf_low, dt, foot_dur))
                T["com"].append(ct(com_lf.copy(), dt, foot_dur))
                T["lh"].append(ct(lh_init, dt, com_dur)), T["rh"].append(ct(rh_init, dt, com_dur))
            elif s == LIFT_DOWN_RF:
                adv = float(np.float32(2) * np.float32(step_length))  # 2 * step_length_ is a float product (walk.cpp:142-143)
                rf_low = tr(rf_low, adv, 0)
                rh_forward = tr(rh_init, adv, 0)
                T["rf"].append(mj(rf_high, rf_low, dt, foot_dur)), T["lf"].append(ct(lf_low, dt, foot_dur))
                T["com"].append(ct(com_lf.copy(), dt, foot_dur))
                T["lh"].append(ct(lh_init, dt, com_dur)), T["rh"].append(mj(rh_init, rh_forward, dt, com_dur))
                rh_init = rh_forward
            elif s == MOVE_COM_RIGHT:
                com_rf = com_rf.copy()
                com_rf[0] = rf_low[0]
                T["rf"].append(ct(rf_low, dt, com_dur)), T["lf"].append(ct(lf_low, dt, com_dur))
                T["com"].append(mj(com_lf, com_rf, dt, com_dur))
                T["lh"].append(ct(lh_init, dt, com_dur)), T["rh"].append(ct(rh_init, dt, com_dur))
            elif s == LIFT_UP_LF:
                lf_high = lf_high.copy()
                lf_high[0] = rf_low[0]
                T["rf"].append(ct(rf_low, dt, foot_dur)), T["lf"].append(mj(lf_low, lf_high, dt, foot_dur))
                T["com"].append(ct(com_rf.copy(), dt, foot_dur))
                T["lh"].append(ct(lh_init, dt, com_dur)), T["rh"].append(ct(rh_init, dt, com_dur))
            elif s == LIFT_DOWN_LF_FINAL:
                lf_low = tr(lf_low, step_length, 0)
                lh_forward = tr(lh_init, step_length, 0)
                T["rf"].append(ct(rf_low, dt, foot_dur)), T["lf"].append(mj(lf_high, lf_low, dt, foot_dur))
                T["com"].append(ct(com_rf.copy(), dt, foot_dur))
                T["lh"].append(mj(lh_init, lh_forward, dt, com_dur)), T["rh"].append(ct(rh_init, dt, com_dur))
            elif s == MOVE_COM_CENTER_FINAL:
                com_init = com_init.copy()
                com_init[:2] = (rf_low[:2] + lf_low[:2]) / 2.0
                T["rf"].append(ct(rf_low, dt, com_dur)), T["lf"].append(ct(lf_low, dt, com_dur))
                T["com"].append(mj(com_rf, com_init, dt, com_dur))
                T["lh"].append(ct(lh_forward, dt, com_dur)), T["rh"].append(ct(rh_forward, dt, com_dur))
        self.T = T
        self.seg_len = [len(x) for x in T["com"]]
        self.total_ticks = int(sum(self.seg_len))
        # play the state machine once (walk.cpp:198-245): refs and mask before the controller update of every tick
        self._refs = np.tile(self.ref0, (self.total_ticks, 1))
        self._mask = np.zeros(self.total_ticks, dtype=np.uint32)
        mask = problem.default_contact_mask
        state, cur, time_ = cycle[0], 0, 0
        for g in range(self.total_ticks):
            if time_ == 0 and state in (LIFT_UP_LF, LF_INIT):
                mask &= ~self.bit["lf"]
            if time_ == 0 and state == LIFT_UP_RF:
                mask &= ~self.bit["rf"]
            if time_ == len(T["com"][cur]) - 1 and state in (LIFT_DOWN_LF, LIFT_DOWN_LF_FINAL):
                mask |= self.bit["lf"]
            if time_ == len(T["com"][cur]) - 1 and state == LIFT_DOWN_RF:
                mask |= self.bit["rf"]
            r = self._refs[g]
            r[self.slot["com"]:self.slot["com"] + 3] = T["com"][cur][time_]
            r[self.slot["com"] + 3:self.slot["com"] + 9] = 0.0  # to_sample(VectorXd): zero derivatives (trajs/utils.hpp:20-25)
            for k in ("lf", "rf", "lh", "rh"):
                o = self.slot[k]
                r[o:o + 3] = T[k][cur][time_]
                r[o + 12:o + 24] = 0.0  # to_sample(SE3): zero derivatives (trajs/utils.hpp:27-34)
            for k in ("lf", "rf"):
                o = self.cslot[k]
                r[o:o + 3] = T[k][cur][time_]
                r[o + 12:o + 24] = 0.0
            self._mask[g] = mask
            time_ += 1
            if time_ == len(T["com"][cur]):
                time_ = 0
                cur += 1
                if cur < len(cycle):
                    state = cycle[cur]

    def tick(self, t):
        """(refs[nref], contact mask) of global tick t; after the last tick the behaviour stops writing (run_ = false)."""
        t = min(int(t), self.total_ticks - 1)
        return self._refs[t], int(self._mask[t])

    def table(self):
        return self._refs, self._mask


class WalkOnSpot:
    """src/behaviors/humanoid/walk_on_spot.cpp:36-208.  INIT is played once, then the six remaining states repeat for ever
    (`_current_traj = ++_current_traj % cycle_.size(); if (_current_traj == 0) _current_traj++`)."""

    INIT, LIFT_UP_LF, LIFT_DOWN_LF, MOVE_COM_LEFT, LIFT_UP_RF, LIFT_DOWN_RF, MOVE_COM_RIGHT = range(7)

    def __init__(self, problem, config):
        c = (config if isinstance(config, dict) else yaml.safe_load(open(config)))["BEHAVIOR"]
        self.P = problem
        dt = problem.struct().dt
        com_dur = float(np.float32(c["traj_com_duration"]))
        foot_dur = float(np.float32(c["traj_foot_duration"]))
        step_height = float(np.float32(c["step_height"]))
        self.ref0 = problem.default_refs()
        self.slot = {k: problem.ref_slot(k)[0] for k in ("com", "lf", "rf")}
        self.cslot = {"lf": problem.ref_slot("contact_lfoot")[0], "rf": problem.ref_slot("contact_rfoot")[0]}
        self.bit = {"lf": 1 << problem.contact_index("contact_lfoot"), "rf": 1 << problem.contact_index("contact_rfoot")}
        pos = lambda n: self.ref0[self.slot[n]:self.slot[n] + 3].copy()  # noqa: E731
        lf_low, rf_low, com_init = pos("lf"), pos("rf"), pos("com")
        lf_high, rf_high = lf_low.copy(), rf_low.copy()
        lf_high[2] += step_height
        rf_high[2] += step_height
        com_lf, com_rf = lf_low.copy(), rf_low.copy()
        com_lf[2] = com_rf[2] = com_init[2]
        mj, ct = min_jerk_vec, constant_traj
        S = self
        self.cycle = [S.INIT, S.LIFT_UP_LF, S.LIFT_DOWN_LF, S.MOVE_COM_LEFT, S.LIFT_UP_RF, S.LIFT_DOWN_RF, S.MOVE_COM_RIGHT]
        self.T = {
            "rf": [ct(rf_low, dt, com_dur), ct(rf_low, dt, foot_dur), ct(rf_low, dt, foot_dur), ct(rf_low, dt, com_dur),
                   mj(rf_low, rf_high, dt, foot_dur), mj(rf_high, rf_low, dt, foot_dur), ct(rf_low, dt, com_dur)],
            "lf": [ct(lf_low, dt, com_dur), mj(lf_low, lf_high, dt, foot_dur), mj(lf_high, lf_low, dt, foot_dur), ct(lf_low, dt, com_dur),
                   ct(lf_low, dt, foot_dur), ct(lf_low, dt, foot_dur), ct(lf_low, dt, com_dur)],
            "com": [mj(com_init, com_rf, dt, com_dur), ct(com_rf, dt, foot_dur), ct(com_rf, dt, foot_dur), mj(com_rf, com_lf, dt, com_dur),
                    ct(com_lf, dt, foot_dur), ct(com_lf, dt, foot_dur), mj(com_lf, com_rf, dt, com_dur)],
        }
        self.seg_len = [len(x) for x in self.T["com"]]

    def play(self, ticks):
        """refs[ticks][nref], masks[ticks]: what the behaviour has written before the controller update of every tick."""
        refs = np.tile(self.ref0, (ticks, 1))
        masks = np.zeros(ticks, dtype=np.uint32)
        mask = self.P.default_contact_mask
        cur, time_ = 0, 0
        state = self.cycle[0]
        S = self
        for g in range(ticks):
            last = len(self.T["com"][cur]) - 1
            if time_ == 0 and state == S.LIFT_UP_LF:
                mask &= ~self.bit["lf"]
            if time_ == 0 and state == S.LIFT_UP_RF:
                mask &= ~self.bit["rf"]
            if time_ == last and state == S.LIFT_DOWN_LF:
                mask |= self.bit["lf"]
            if time_ == last and state == S.LIFT_DOWN_RF:
                mask |= self.bit["rf"]
            r = refs[g]
            r[self.slot["com"]:self.slot["com"] + 3] = self.T["com"][cur][time_]
            r[self.slot["com"] + 3:self.slot["com"] + 9] = 0.0
            for k in ("lf", "rf"):
                for o in (self.slot[k], self.cslot[k]):
                    r[o:o + 3] = self.T[k][cur][time_]
                    r[o + 12:o + 24] = 0.0
            masks[g] = mask
            time_ += 1
            if time_ == len(self.T["com"][cur]):
                time_ = 0
                cur = (cur + 1) % len(self.cycle)
                if cur == 0:
                    cur = 1
                state = self.cycle[cur]
        return refs, masks
