"""safety_oracle.py -- CPU restatement of the controller's command gating.  TEST INFRASTRUCTURE ONLY.

Follows src/controllers/controller.cpp of the reference line by line for ONE robot:
  _solve():  :238-242 snapshot while send_cmd_;  :244-282 solve, q_solver_ always, q_ / tau_ / dq_ / ddq_ only while send_cmd_;
             :284-308 a non-optimal QP throws (nothing below runs);  :309-312 model collision check on this tick's data, latch
  qp_step_back(): :445-450 with include/inria_wbc/controllers/controller.hpp:170-171
The cycle itself and the sphere test are the C++ oracle's (tests/oracle_lib.py)."""
import numpy as np


class SafeController:
    def __init__(self, ol, problem, q, v, spheres=None):
        """spheres: (p_radius, joint_member, nmembers) of the collision model, or None (check_model_collisions: false)"""
        self.ol, self.P = ol, problem
        self.q_tsid, self.v_tsid = np.array(q, float), np.array(v, float)
        nv, na = problem.nv, problem.na
        self.send_cmd, self.is_model_colliding = True, False
        self.q_prev, self.v_prev = np.zeros(problem.nq), np.zeros(nv)
        self.q_, self.q_solver, self.tau_, self.dq_, self.ddq_ = np.zeros(nv), np.zeros(nv), np.zeros(na), np.zeros(nv), np.zeros(nv)
        self.spheres = spheres

    def solve(self, refs, contact_mask=None):
        q, dq = self.q_tsid.copy(), self.v_tsid.copy()
        if self.send_cmd:                                                   # :238-242
            self.q_prev, self.v_prev = q.copy(), dq.copy()
        o = self.ol.cycle_batch(self.P, q[None], dq[None], refs, contact_mask=contact_mask)
        if o["status"][0] != 0:                                             # :284-308
            raise RuntimeError("Controller failed, can't solve problem. Status : %d" % o["status"][0])
        self.v_tsid, self.q_tsid = o["v_next"][0], o["q_next"][0]           # :254-255
        self.q_solver = o["q_cmd"][0].copy()                                # :266 / :273
        if self.send_cmd:                                                   # :267-282
            self.q_, self.tau_ = o["q_cmd"][0].copy(), o["tau"][0].copy()
            self.dq_, self.ddq_ = o["v_next"][0].copy(), o["dv"][0].copy()
        if self.spheres is not None:                                        # :309-310 (tsid's data: the state this tick started from)
            pr, jm, nm = self.spheres
            self.is_model_colliding = bool(self.ol.collision(self.P, q, pr, jm, nm)[0])
        if self.is_model_colliding and not np.all(np.abs(self.q_) <= 1e-3):  # :311-312  !q_.isZero(1e-3)
            self.send_cmd = False

    def qp_step_back(self):                                                 # :445-450
        self.q_tsid, self.v_tsid = self.q_prev.copy(), self.v_prev.copy()
