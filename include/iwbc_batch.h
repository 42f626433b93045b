/* iwbc_batch.h -- C ABI of the B200-native batched whole-body-control cycle (libiwbc_b200.so).
 *
 * Drop-in boundary for ONE path of resibots/inria_wbc: `Controller::update` -> `_solve`
 * (pinocchio kinematics/dynamics -> TSID task/contact stacking -> eiquadprog QP -> decode ->
 * integrate), for thousands to millions of independent robot states per call.  Plain pointers
 * and sizes only; every entry point names the reference interface it replaces.  Errors: functions
 * return 0 on success and a negative code otherwise; `iwbc_last_error()` gives the message (the
 * reference throws `inria_wbc::Exception`, include/inria_wbc/exceptions.hpp:52-94).  A QP that is
 * not solved to optimality is *data* (`status[i]`), not an error: the reference throws there
 * (src/controllers/controller.cpp:284-308), a batch cannot.
 *
 * There is no CPU fallback: the compute entry points fail with IWBC_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef IWBC_BATCH_H
#define IWBC_BATCH_H

#include "iwbc_problem.h"

#ifdef __cplusplus
extern "C" {
#endif

#define IWBC_OK 0
#define IWBC_ERR_CONFIG -1   /* YAML / URDF / SRDF problem (IWBC_CHECK / IWBC_ASSERT in the reference) */
#define IWBC_ERR_ARG -2      /* bad argument (size asserts of Controller::update, controller.cpp:163-205) */
#define IWBC_ERR_CUDA -3     /* CUDA runtime error or no usable device */
#define IWBC_ERR_UNSUPPORTED -4

#define IWBC_ACTIVE_WORDS 12 /* 32-bit words of the per-instance active-set bitmask (>= mi bits) */

typedef struct iwbc_problem iwbc_problem; /* compiled controller: model + task stack (host) */
typedef struct iwbc_batch iwbc_batch;     /* one GPU's batch workspace bound to a problem   */

const char* iwbc_last_error(void);
int iwbc_abi_version(void);

/* ------------------------------------------------------------------ construction (host only)
 * Replaces `controllers::Factory::instance().create(name, YAML::LoadFile(path))`
 * (include/inria_wbc/utils/factory.hpp:51-63, include/inria_wbc/controllers/controller.hpp:235-237)
 * and everything the constructors chain does (src/controllers/controller.cpp:65-130,
 * src/controllers/pos_tracker.cpp:40-159, src/controllers/humanoid_pos_tracker.cpp:37-79,
 * src/controllers/talos_pos_tracker.cpp:37-60).  The controller name is read from CONTROLLER.name
 * ("pos-tracker", "humanoid-pos-tracker", "talos-pos-tracker").
 *   controller_yaml : path of the controller file (e.g. etc/talos/talos_pos_tracker.yaml)
 *   robots_path     : directory `CONTROLLER.urdf` is resolved against (utheque's role)
 *   etc_path        : replaces a leading "@ETC@" of CONTROLLER.base_path (NULL: dirname/.. of the yaml)
 *   overrides_yaml  : optional YAML text of CONTROLLER keys to override, the way the reference's
 *                     tests mutate the loaded node (tests/test_determinism.cpp:66-70)
 */
iwbc_problem* iwbc_problem_load(const char* controller_yaml, const char* robots_path, const char* etc_path,
    const char* overrides_yaml);
void iwbc_problem_free(iwbc_problem* p);

/* flat description shared with the device (and with the CPU oracle in tests) */
const iwbc_problem_t* iwbc_problem_data(const iwbc_problem* p);

/* sizes: nq nv na ncontacts nref, and the QP shape (n, me, mi) for a given contact mask
 * (PosTracker prints these: src/controllers/pos_tracker.cpp:150-158) */
int iwbc_problem_dims(const iwbc_problem* p, int* nq, int* nv, int* na, int* ncontacts, int* nref);
int iwbc_problem_qp_shape(const iwbc_problem* p, unsigned contact_mask, int* n, int* me, int* mi, int* cost_rows);
unsigned iwbc_problem_default_contact_mask(const iwbc_problem* p);

/* construction-time references: what make_se3 / make_com / make_posture / make_contact_task store
 * (src/controllers/tasks.cpp:74-77,108,217,361-362) + init_com.  out has nref doubles. */
int iwbc_problem_default_refs(const iwbc_problem* p, double* out);
/* offset/length of a task's or contact's block inside the reference vector: the batched form of
 * PosTracker::set_se3_ref / set_com_ref / set_momentum_ref / set_contact_se3_ref / set_posture_ref
 * (include/inria_wbc/controllers/pos_tracker.hpp:52-58) is "write into refs[i][offset..]" */
int iwbc_problem_ref_slot(const iwbc_problem* p, const char* task_or_contact, int* offset, int* length);
/* PosTracker::set_task_weight / Behavior::_customize_tasks (pos_tracker.hpp:67, src/behaviors/behavior.cpp:7-22) */
int iwbc_problem_set_task_weight(iwbc_problem* p, const char* task, double weight);
int iwbc_problem_contact_index(const iwbc_problem* p, const char* contact);
/* names: kind 0 = actuated dofs (Controller::controllable_dofs), 1 = tasks, 2 = contacts, 3 = mimic dofs.
 * Returns the count; name i is copied to buf (if buf != NULL). */
int iwbc_problem_name(const iwbc_problem* p, int kind, int i, char* buf, int buflen);
/* CONTROLLER.closed_loop (src/controllers/controller.cpp:86): 1 when the controller expects the state from the sensors */
int iwbc_problem_closed_loop(const iwbc_problem* p);
/* reference configuration q0 (nq doubles, pinocchio layout, quaternion normalised: pos_tracker.cpp:62-77) */
int iwbc_problem_q0(const iwbc_problem* p, double* out);

/* ------------------------------------------------------------------ batched update (GPU)
 * Instance-major arrays: q[B][nq], v[B][nv], refs[B][nref] (refs_stride = nref) or one shared
 * reference vector (refs_stride = 0); contact_mask[B] or NULL (all default-active contacts).
 * Instances with different masks are bucketed inside the call.
 */
typedef struct {
    double* tau;       /* [B][na]   getActuatorForces            (controller.cpp:250)          */
    double* dv;        /* [B][nv]   getAccelerations             (controller.cpp:251)          */
    double* f;         /* [B][12*ncontacts] getContactForces, 0 for inactive (controller.cpp:258-261) */
    double* q_next;    /* [B][nq]   pinocchio::integrate(q, dt * v_next)  (controller.cpp:255) */
    double* v_next;    /* [B][nv]   dq + dt * dv                 (controller.cpp:254)          */
    double* q_cmd;     /* [B][nv]   "DART" layout q_: base angle-axis (controller.cpp:263-276) */
    double* momentum;  /* [B][3]    Ag.bottomRows(3) * dq        (controller.cpp:245)          */
    int* status;       /* [B]       iwbc_status (HQP status)     (controller.cpp:249,287-304)  */
    int* iters;        /* [B]       active-set iterations                                      */
    unsigned* active;  /* [B][IWBC_ACTIVE_WORDS] bitmask over the instance's CI rows           */
} iwbc_results_t;      /* any pointer may be NULL */

iwbc_batch* iwbc_batch_create(const iwbc_problem* p, int device, int max_batch);
void iwbc_batch_free(iwbc_batch* b);

/* `Controller::update(sensor_data)` for B states held in HOST memory (closed-loop form: q, v are
 * the sensor-side state, controller.cpp:163-205).  Copies in, runs the cycle, copies the requested
 * results out and synchronises.  Pinned host buffers make the copies asynchronous. */
int iwbc_batch_update(iwbc_batch* b, int B, const double* q, const double* v, const double* refs, long refs_stride,
    const unsigned* contact_mask, const iwbc_results_t* out);

/* Same, all pointers are DEVICE pointers on the batch's device; enqueued on `stream`
 * (a cudaStream_t passed as void*), no synchronisation.  Uniform contact mask per call. */
int iwbc_batch_update_device(iwbc_batch* b, int B, const double* q, const double* v, const double* refs, long refs_stride,
    unsigned contact_mask, const iwbc_results_t* out, void* stream);

/* Open-loop ticks (closed_loop: false, Controller::_solve() on the integrated state): advances
 * the device-resident state (q, v) `ticks` times in place with constant references.  An instance whose QP is not solved to
 * optimality at some tick is frozen at that state (controller.cpp:249-256 throws before integrating) and keeps that status. */
int iwbc_batch_rollout_device(iwbc_batch* b, int B, double* q, double* v, const double* refs, long refs_stride,
    unsigned contact_mask, int ticks, const iwbc_results_t* out_last, void* stream);

/* number of kernel launches issued by this batch so far (bench.py's gpu_launches) */
long iwbc_batch_launch_count(const iwbc_batch* b);
/* average device time (ms) of the cycle kernel over the launches since the last reset, measured
 * with CUDA events on the launching stream; reset = 1 clears the accumulators */
double iwbc_batch_kernel_ms(iwbc_batch* b, int reset);
int iwbc_batch_enable_timing(iwbc_batch* b, int on);

/* clock64() timestamps of the phases of the last iwbc_batch_debug_instance call (instance 0):
 * start, after k_dyn, after k_asm, after factor+inverse, after equalities, after the active-set loop, after decode */
int iwbc_batch_debug_cycles(const iwbc_batch* b, long long* out7);
/* steady-state phase profile: while enabled, every instance adds its per-phase SM-cycle counts (clock64 deltas of the
 * group's first thread) to device accumulators.  The call synchronises the device, copies the accumulators to out7 (if
 * not NULL, 16 entries: [0..5] summed cycles of dyn, asm, factor, equalities, active-set loop, decode of the
 * active-set kernel and [6] its instance count; [8..13] dyn, rows, H + Cholesky, QR + solve, termination test, decode of
 * the fast kernel and [14] the instances it finished), clears them and switches profiling on/off.  Development aid; not part of the reference surface. */
int iwbc_batch_phase_profile(iwbc_batch* b, int enable, unsigned long long* out16);
/* phases of the active-set loop of the continuation kernel accumulated up to the last iwbc_batch_phase_profile call (SM cycles as
 * seen by thread 0, barrier waits included): out8 = { pass (s, psi, pick), n+ load, d, z + r + ratio test + coefficients,
 * step + rotations, drops, number of step directions, number of added constraints, and inside the z phase the time to finish
 * of: the last warp (r = R^-1 d, ratio test), the coefficient warp, warp 0's z rows; 5 spare }.  out16 has 16 entries.  Development aid. */
int iwbc_batch_loop_profile(const iwbc_batch* b, unsigned long long* out16);
/* launch shape of the cycle kernel for the last contact mask: threads per instance, instances per CTA, grid, dynamic smem */
int iwbc_batch_launch_info(const iwbc_batch* b, int* G, int* ipc, int* grid, long* smem_bytes);

/* launch shape of the fast pass / active-set continuation pair (floating-base robots with contacts):
 * out6 = { enabled, fast grid, fast dynamic smem bytes, continuation grid, continuation smem bytes, threads per CTA } */
int iwbc_batch_fast_info(const iwbc_batch* b, long* out6);

/* measured FP64 FMA throughput of the device in TFLOP/s (dependent-FMA chains, 8-way ILP): the roofline
 * denominator of this FP64 CUDA-core-bound path; < 0 on error */
double iwbc_measure_fp64_tflops(int device);

/* stage dump of ONE instance for parity tests (device results copied to host; any pointer may be NULL):
 * M[nv*nv] nle[nv] H[n*n] g[n] x[n], all row-major */
/* Controller::cost(task) (controller.hpp:148-152: norm of the task's own constraint residual A ddq - b at the accelerations `ddq`,
 * usually the last solution) and PosTracker::objective_value() (pos_tracker.hpp:44: eiquadprog's objective x'Hx/2 + g'x at the
 * solution) of ONE robot: logging getters of the reference, re-evaluated on demand from a stage dump of that robot at the state the
 * QP was solved at (the batched cycle does not carry H or the task rows off the SM).  `task`: a level-1 task name (se3, com,
 * momentum, self-collision, posture; `torque`: tau[na] of that solve must be passed, NULL otherwise); either output may be NULL. */
int iwbc_batch_task_cost(iwbc_batch* b, const double* q, const double* v, const double* ref, unsigned contact_mask, const char* task,
    const double* ddq, const double* tau, double* cost);
int iwbc_batch_objective_value(iwbc_batch* b, const double* q, const double* v, const double* ref, unsigned contact_mask, double* objective);

int iwbc_batch_debug_instance(iwbc_batch* b, const double* q, const double* v, const double* ref, unsigned contact_mask,
    double* M, double* nle, double* H, double* g, double* x);

/* ------------------------------------------------------------------ behaviours on the device
 * Replaces `behaviors::Factory::instance().create(name, controller, YAML::LoadFile(path))`
 * (include/inria_wbc/behaviors/behavior.hpp:28-30) and the per-tick `Behavior::update` setters
 * (e.g. src/behaviors/humanoid/walk.cpp:198-246): the behaviour's waypoint tables
 * (include/inria_wbc/trajs/trajectory_generator.hpp:29-160) are compiled into a segment table that
 * the GPU evaluates in closed form for every (instance, tick).  BEHAVIOR.name selects the
 * As in the reference, constructing a behaviour applies BEHAVIOR.customize_task_weights to the controller
 * (Behavior::_customize_tasks, src/behaviors/behavior.cpp:7-22), which is why `p` is not const.  Available
 * behaviour: "humanoid::walk" (src/behaviors/humanoid/walk.cpp), "humanoid::walk-on-spot"
 * (src/behaviors/humanoid/walk_on_spot.cpp), "humanoid::clapping" (src/behaviors/humanoid/clapping.cpp), "humanoid::move-feet" (src/behaviors/humanoid/move_feet.cpp), "generic::cartesian_traj" (src/behaviors/generic/cartesian_traj.cpp), "generic::cartesian"
 * (src/behaviors/generic/cartesian.cpp:8-94), "humanoid::move_com" (src/behaviors/humanoid/move_com.cpp:8-61); overrides_yaml: optional YAML text of BEHAVIOR keys to override.
 */
typedef struct iwbc_behavior iwbc_behavior;
iwbc_behavior* iwbc_behavior_load(iwbc_problem* p, const char* behavior_yaml, const char* overrides_yaml);
void iwbc_behavior_free(iwbc_behavior* h);
/* length in control ticks, segment / channel counts, the distinct contact masks (<= 8, most frequent first) */
int iwbc_behavior_info(const iwbc_behavior* h, int* total_ticks, int* nsegments, int* nchannels, int* nmasks, unsigned* masks8);
/* one segment of the table (host): first tick, length, contact mask during / at the last tick of the segment, and per
 * channel 8 doubles {p0(3), p1(3), duration, minjerk ? 1 : 0} (waypoints may be NULL) */
int iwbc_behavior_segment(const iwbc_behavior* h, int seg, int* start, int* ticks, unsigned* mask, unsigned* mask_last, double* waypoints);
/* DEVICE pointers.  refs[B][nref] and contact_mask[B] (may be NULL) as the behaviour leaves them before the controller
 * update of global tick `tick + phase[i]` (phase: per-instance offset or NULL); init_refs != 0 first fills every row
 * with the construction-time references.  Enqueued on `stream`, no synchronisation. */
int iwbc_behavior_eval_device(iwbc_behavior* h, int device, int B, const int* phase, int tick, double* refs, unsigned* contact_mask,
    int init_refs, void* stream);
/* Closed device-side roll-out: `ticks` times { behaviour -> refs, masks; instances bucketed by contact mask on the device;
 * one shape-uniform launch set of the control cycle per mask; (q, v) integrated in place }.  No host round trip and no
 * reference upload per tick.  refs[B][nref] is caller-owned device scratch (kept between calls so that a roll-out can be
 * continued with init_refs = 0); out_last (device pointers, any may be NULL): results of the last tick. */
int iwbc_batch_behavior_rollout_device(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, double* refs, const int* phase,
    int tick0, int ticks, int init_refs, const iwbc_results_t* out_last, unsigned* contact_mask, void* stream);

/* The same roll-out with HOST buffers: q[B][nq], v[B][nv] are uploaded, advanced `ticks` ticks on the device and written back; phase[B]
 * (host) or NULL; out_last / contact_mask are host arrays (any may be NULL).  tick0 == 0 starts from the construction-time references; a
 * later call (tick0 > 0, same B) continues where the previous one stopped (the reference rows live in the batch's own device
 * workspace: an iwbc_batch_update call in between overwrites them -- restart with tick0 = 0 semantics or use the device form).  Synchronous. */
int iwbc_batch_behavior_run(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, const int* phase, int tick0, int ticks,
    const iwbc_results_t* out_last, unsigned* contact_mask);

/* The two roll-outs with a per-tick command trace: q_cmd_trace[ticks][B][nv] receives the "DART"-layout command vector of EVERY
 * tick, i.e. what the reference's tests record with `controller->q(false)` after each `behavior->update()`
 * (tests/test_determinism.cpp:35-40, tests/test_franka.cpp:175-179).  Device form: a device buffer; host form: a host buffer
 * (drained chunk by chunk from a device ring).  A tick whose QP is not solved to optimality does not advance that instance's state
 * (the reference throws before q_tsid_ / v_tsid_ are touched, src/controllers/controller.cpp:249-256) and its status is kept for
 * the rest of the call; the same holds for every roll-out entry point above. */
int iwbc_batch_behavior_rollout_trace_device(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, double* refs, const int* phase,
    int tick0, int ticks, int init_refs, const iwbc_results_t* out_last, unsigned* contact_mask, double* q_cmd_trace, void* stream);
int iwbc_batch_behavior_run_trace(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, const int* phase, int tick0, int ticks,
    const iwbc_results_t* out_last, unsigned* contact_mask, double* q_cmd_trace);

/* ------------------------------------------------------------------ post-solve model collision check
 * Replaces `utils::CollisionCheck::load_collision_file` + `is_colliding(model, data)` (src/safety/collision_check.cpp:5-87),
 * which the controller runs after every solve when CONTROLLER.check_model_collisions is set (src/controllers/controller.cpp:89-98,
 * 309-310): links covered by spheres {x, y, z, diameter} grouped in members; two spheres of DIFFERENT members closer than the sum of
 * their radii = collision.  collision_yaml: the file CONTROLLER.collision_path names (e.g. etc/talos/collisions/talos_collisions_margin.yaml).
 */
typedef struct iwbc_collision iwbc_collision;
iwbc_collision* iwbc_collision_load(const iwbc_problem* p, const char* collision_yaml);
void iwbc_collision_free(iwbc_collision* h);
/* sphere / member counts; links named by the file that the model does not have (the reference skips them silently) */
int iwbc_collision_info(const iwbc_collision* h, int* nspheres, int* nmembers, int* skipped_links);
/* the compiled table (host): per sphere {x, y, z, radius} and {supporting joint (-1 = universe), member index (members sorted by name)} */
int iwbc_collision_spheres(const iwbc_collision* h, double* p_radius4, int* joint_member2);
/* DEVICE pointers: q[B][nq] (the state the tick started from: the reference checks tsid's data), colliding[B] (0 / 1, the
 * batched is_model_colliding_), pair4[B][4] or NULL = {member a, sphere i, member b, sphere j} of the first hit in the reference's
 * iteration order (-1 when free).  Enqueued on `stream` (NULL: the batch's own stream), no synchronisation. */
int iwbc_batch_check_collisions_device(iwbc_batch* b, iwbc_collision* h, int B, const double* q, int* colliding, int* pair4, void* stream);

/* ------------------------------------------------------------------ send_cmd latch and qp_step_back (second half)
 * Per-instance form of the controller's command gating (src/controllers/controller.cpp):
 *   :238-242  while send_cmd_ the state a tick starts from is remembered (q_tsid_prev_, v_tsid_prev_);
 *   :263-282  q_solver_ follows every OPTIMAL solve, the command vectors q_ / tau_ / dq_ / ddq_ only while send_cmd_;
 *   :309-312  after the solve, a model collision (on the state the tick started from) latches send_cmd_ = false unless the
 *             command is still all zero (`!q_.isZero(1e-3)`); the controller's own integrated state keeps advancing;
 *   :445-450  qp_step_back() puts the remembered state back (include/inria_wbc/controllers/controller.hpp:170-171);
 *   set_send_cmd(true) (controller.hpp:167) = write 1 into send_cmd[i].
 * All pointers are DEVICE pointers; the caller owns the arrays and initialises send_cmd to 1 and q_cmd to 0 (q_ starts at zero).
 */
typedef struct {
    int* send_cmd;     /* [B]      in/out: Controller::send_cmd_                                  */
    int* colliding;    /* [B]      out: is_model_colliding_ of the last tick (needed iff a collision model is given) */
    double* q_prev;    /* [B][nq]  q_tsid_prev_                                                   */
    double* v_prev;    /* [B][nv]  v_tsid_prev_                                                   */
    double* q_solver;  /* [B][nv]  q_solver_ ("DART" layout) -- may be NULL                        */
    double* q_cmd;     /* [B][nv]  q_                                                             */
    double* tau_cmd;   /* [B][na]  tau_ without its six leading zeros -- may be NULL               */
    double* dq_cmd;    /* [B][nv]  dq_  -- may be NULL                                            */
    double* ddq_cmd;   /* [B][nv]  ddq_ -- may be NULL                                            */
} iwbc_safety_t;
/* `ticks` open-loop ticks on the device-resident state (q, v) in place with the gating above; collision may be NULL
 * (CONTROLLER.check_model_collisions: false); status[B] (may be NULL) as in iwbc_batch_rollout_device. */
int iwbc_batch_safe_rollout_device(iwbc_batch* b, iwbc_collision* collision, int B, double* q, double* v, const double* refs, long refs_stride,
    unsigned contact_mask, int ticks, const iwbc_safety_t* safety, int* status, void* stream);
/* qp_step_back(): (q, v) <- (q_prev, v_prev) for the instances with select[i] != 0 (select == NULL: all) */
int iwbc_batch_qp_step_back_device(iwbc_batch* b, int B, double* q, double* v, const iwbc_safety_t* safety, const int* select, void* stream);

/* ------------------------------------------------------------------ stabilizer: the sensor-driven pre-step
 * `HumanoidPosTracker::update(sensor_data)` with CONTROLLER.stabilizer.activated (src/controllers/humanoid_pos_tracker.cpp:134-302):
 * CoP estimate from the feet's force / torque sensors (src/estimators/cop.cpp:10-96) through moving-average filters
 * (include/inria_wbc/estimators/filtering.hpp:17-146), then the CoM / ZMP-distributor / ankle / foot-force-difference admittances
 * (src/stabilizers/stabilizer.cpp:31-59, 277-311 + 340-429, 168-224, 226-275) rewrite the references of ONE solve; for
 * talos-pos-tracker the IMU momentum admittance comes first (src/controllers/talos_pos_tracker.cpp:135-150).  The parameter sets are
 * the reference's stab_*.yaml files named by CONTROLLER.stabilizer.params_{fixed_base,ss,ds}, read unchanged at load time.
 * A controller loaded with stabilizer.activated = true REFUSES the plain update / roll-out entry points ("the stabilizer needs the
 * LF force", :180) -- load it with `stabilizer: {activated: false}` for open-loop use, as tests/test_determinism.cpp:67 does.
 *
 * iwbc_stabilizer = the per-robot state that survives a tick, for the batch's max_batch robots: filter windows, normal forces of the
 * last solve, the contacts' force references and installed regularisation weights, the state tsid's data was computed on.
 *   sensors[B][15]        : lf_force(3) rf_force(3) lf_torque(3) rf_torque(3) imu_vel(3)   (SensorData keys of the same names)
 *   contact_mask[B]       : activated contacts per robot (NULL: the default mask); both feet, left only or right only
 *   behavior_type[B]      : iwbc_behavior_type per robot = Controller::set_behavior_type (NULL: behavior_type_all for everybody)
 *   refs / refs_stride    : the behaviours' reference rows; they are NOT modified (the reference puts its references back after the
 *                           solve, :291-301); refs_used[B][nref] (may be NULL) receives the rows the solve saw
 */
typedef struct iwbc_stabilizer iwbc_stabilizer;
iwbc_stabilizer* iwbc_stabilizer_create(iwbc_batch* b);
void iwbc_stabilizer_free(iwbc_stabilizer* s);
int iwbc_stabilizer_reset(iwbc_stabilizer* s, void* stream); /* fresh controllers: empty windows, data state = (q0, 0) */
/* DEVICE pointers, enqueued on `stream`, no synchronisation */
int iwbc_batch_update_sensors_device(iwbc_batch* b, iwbc_stabilizer* s, int B, const double* q, const double* v, const double* refs, long refs_stride,
    const unsigned* contact_mask, const int* behavior_type, int behavior_type_all, const double* sensors, const iwbc_results_t* out, double* refs_used,
    void* stream);
/* HOST pointers; synchronous */
int iwbc_batch_update_sensors(iwbc_batch* b, iwbc_stabilizer* s, int B, const double* q, const double* v, const double* refs, long refs_stride,
    const unsigned* contact_mask, const int* behavior_type, int behavior_type_all, const double* sensors, const iwbc_results_t* out, double* refs_used);

/* ------------------------------------------------------------------ one call, several GPUs (SURVEY.md 7 "shard/gather", 8(e))
 * The batched form of `Controller::update` for a batch that spans the GPUs of one node: contiguous shards
 * [g B / G, (g + 1) B / G), one iwbc_batch and one worker thread per device, every shard read from / written into its own slice
 * of the caller's HOST arrays (the "host gather of torques and accelerations" of the north star); no collective.  Allocate the
 * arrays with iwbc_host_alloc (pinned, portable) for the copies to run at PCIe speed on all devices at once.
 *   devices: ndev device ordinals (NULL: 0 .. ndev-1; an ordinal may repeat: several shards on one GPU) */
typedef struct iwbc_multi iwbc_multi;
iwbc_multi* iwbc_multi_create(const iwbc_problem* p, int ndev, const int* devices, int max_batch_total);
void iwbc_multi_free(iwbc_multi* m);
int iwbc_multi_device_count(const iwbc_multi* m);
int iwbc_multi_update(iwbc_multi* m, int B, const double* q, const double* v, const double* refs, long refs_stride,
    const unsigned* contact_mask, const iwbc_results_t* out);
long iwbc_multi_launch_count(const iwbc_multi* m);
/* shard g of nshards of a batch of B: first instance and count (host only) */
int iwbc_shard_range(int B, int nshards, int g, int* first, int* count);
void* iwbc_host_alloc(size_t bytes);
void iwbc_host_free(void* p);

/* ------------------------------------------------------------------ solver interface
 * Batched form of `tsid::solvers::SolverHQPBase::solve` for the EiquadprogFast backend
 * (src/controllers/pos_tracker.cpp:96,102; src/controllers/controller.cpp:247): B dense QPs of one
 * shape, HOST pointers, row-major:  min 1/2 x'Hx + g'x  s.t. CE x + ce0 = 0, CI x + ci0 >= 0.
 * status[i]: iwbc_status.  active[i][IWBC_ACTIVE_WORDS]: final active inequality set. */
int iwbc_qp_solve_batch(int device, int B, int n, int me, int mi, const double* H, const double* g, const double* CE,
    const double* ce0, const double* CI, const double* ci0, double* x, int* status, int* iters, unsigned* active);

#ifdef __cplusplus
}
#endif
#endif /* IWBC_BATCH_H */
