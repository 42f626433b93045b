/* iwbc_problem.h -- flat, pointer-free description of one compiled whole-body-control problem.
 *
 * One `iwbc_problem_t` = what the reference builds once at construction time
 * (/root/reference/src/controllers/controller.cpp:65-130 model load,
 *  /root/reference/src/controllers/pos_tracker.cpp:40-159 TSID stack,
 *  /root/reference/src/controllers/tasks.cpp:38-404 task constants):
 * the kinematic tree, the frames the tasks use, the task / contact tables and the order in which
 * they enter the two-level HQP.  It is constant per (robot, task file); only the per-instance state
 * (q, v), reference block and contact mask change from tick to tick.
 *
 * The same bytes are (a) uploaded to the GPU (`__constant__` header + global arrays) by the product
 * library and (b) handed to the CPU oracle (oracle/), so both sides provably solve the same problem.
 * Plain C, fixed-size arrays, no pointers: it can be memcpy'd, hashed and serialised as is.
 */
#ifndef IWBC_PROBLEM_H
#define IWBC_PROBLEM_H

#ifdef __cplusplus
extern "C" {
#endif

#define IWBC_MAXJ 56   /* movable joints incl. the free-flyer (Talos: 45) */
#define IWBC_MAXNV 64  /* velocity dimension (Talos: 50) */
#define IWBC_MAXF 48   /* frames referenced by tasks / contacts */
#define IWBC_MAXT 32   /* level-1 and level-0 motion / actuation tasks */
#define IWBC_MAXC 4    /* rigid 6-D contacts */
#define IWBC_MAXAV 12  /* avoided frames per self-collision task */
#define IWBC_MAXL0 16  /* level-0 constraint blocks */
#define IWBC_MAXL1 40  /* level-1 cost blocks */
#define IWBC_MAXWV 4   /* distinct force-regularisation weight vectors (Contact6d::setRegularizationTaskWeightVector values in use) */

enum iwbc_joint_type { IWBC_J_FREEFLYER = 0, IWBC_J_REVOLUTE = 1, IWBC_J_PRISMATIC = 2 };

enum iwbc_task_type {
    IWBC_T_SE3 = 0,       /* TaskSE3Equality (in-tree copy: example_project/src/tsid/ex_task.cpp:175-247) */
    IWBC_T_COM = 1,       /* TaskComEquality */
    IWBC_T_POSTURE = 2,   /* TaskJointPosture */
    IWBC_T_MOMENTUM = 3,  /* TaskMEquality (src/tsid/task-momentum-equality.cpp:144-174) */
    IWBC_T_SELFCOL = 4,   /* TaskSelfCollision (src/tsid/task-self-collision.cpp:84-204) */
    IWBC_T_BOUNDS = 5,    /* TaskJointPosVelAccBounds */
    IWBC_T_ACTBOUNDS = 6, /* TaskActuationBounds */
    IWBC_T_COP = 8,       /* TaskCopEquality through addForceTask (tasks.cpp:156-178): level-1 cost on the contact forces,
                             sum over the active contacts' points j of (p_j - cop_ref) (n' f_j) = 0 (p_j: world position of the contact
                             point, f_j: its 3 force coefficients); reference block = [cop_ref (3) ; normal (3)] (setReference /
                             setContactNormal; the reference's factory sets (0,0,0), TSID's default normal is (0,0,1)).  Upstream TSID
                             is not in the reference tree: restated from the published task-cop-equality.cpp and the m_taskContactForces
                             loop of inverse-dynamics-formulation-acc-force.cpp ("by default associated to all contact forces"). */
    IWBC_T_TORQUE = 7     /* TaskActuationEquality through addActuationTask (tasks.cpp:227-271): level-1 cost on the joint torques
                             tau = M_a dv + h_a - J_a^T f; reference block = [tau_ref (na) ; weight vector (na)] (setReference /
                             setWeightVector); mask: na entries.  Upstream TSID source is not in the reference tree: restated
                             from the published task-actuation-equality.cpp / inverse-dynamics-formulation-acc-force.cpp. */
};

/* kinds of level-0 / level-1 blocks, in HQPData insertion order */
enum iwbc_block_kind {
    IWBC_B_BASEDYN = 0,   /* [M_u | -J_u^T] x = -h_u                    (level 0, equality)   */
    IWBC_B_CONTACT_MOTION = 1, /* contact c: J_c dv = a_des - drift      (level 0, equality)   */
    IWBC_B_CONTACT_FORCE = 2,  /* contact c: lb <= B f_c <= ub           (level 0, inequality) */
    IWBC_B_CONTACT_REG = 3,    /* contact c: diag(w) T f_c = diag(w) f_ref (level 1, cost)     */
    IWBC_B_TASK = 4            /* task t (equality cost at level 1, inequality at level 0)     */
};

typedef struct {
    int type;            /* iwbc_task_type */
    int level;           /* 0 = constraint, 1 = cost */
    int nrows;           /* rows after masking */
    int frame;           /* SE3 / SELFCOL: index into the frame table */
    int ref_off;         /* offset (in doubles) of this task's block in the per-instance reference vector, -1 = none */
    int nav;             /* SELFCOL: number of avoided frames */
    int av_frame[IWBC_MAXAV];
    unsigned char mask[IWBC_MAXNV]; /* SE3/MOMENTUM: 6, COM: 3, POSTURE: na entries */
    double weight;
    double kp[6], kd[6]; /* SE3/COM/MOMENTUM per-axis gains; POSTURE/SELFCOL use kp[0], kd[0] */
    double av_r0[IWBC_MAXAV];
    double radius;       /* SELFCOL */
    double sc_k, sc_sp, sc_m; /* SELFCOL 5PL constants k, s_p, m (task-self-collision.cpp:163-165) */
} iwbc_task_t;

typedef struct {
    int frame;           /* index into the frame table (the contact joint's frame) */
    int ref_off;         /* offset of the contact's SE3 reference block */
    int active_default;  /* in contact at construction */
    int pad_;
    double kp[6], kd[6];
    double T[6 * 12];    /* force generator, row-major 6x12 */
    double B[17 * 12];   /* friction pyramid + normal-force row, row-major 17x12 */
    double lb[17], ub[17];
    double wreg[6];      /* force-regularisation row weights (1,1,1e-3,2,2,2) */
    double fref[6];
    double w_forcereg;   /* level-1 weight of the regularisation block (1e-3) */
} iwbc_contact_t;

typedef struct {
    int kind;            /* iwbc_block_kind */
    int index;           /* task or contact index */
} iwbc_block_t;

/* one stabilizer parameter set = StabConfig (include/inria_wbc/stabilizers/stabilizer_conf.hpp:14-92), one per behaviour type */
enum iwbc_behavior_type { IWBC_BT_FIXED_BASE = 0, IWBC_BT_SINGLE_SUPPORT = 1, IWBC_BT_DOUBLE_SUPPORT = 2 };
typedef struct {
    int filter_size;
    int use_zmp;
    int use_momentum;
    int zmp_w_variant;   /* index of zmp_w in wreg_variants */
    double com_gains[6], ankle_gains[6], ffda_gains[3], pad_;
    double zmp_p[6], zmp_d[6], zmp_w[6], momentum_p[6], momentum_d[6];
} iwbc_stab_conf_t;

typedef struct {
    unsigned int magic;  /* 'IWBC' */
    unsigned int version;
    /* ---- sizes */
    int nj;              /* movable joints (free-flyer counts as one) */
    int nq, nv, na;      /* configuration, velocity, actuated dimensions */
    int floating_base;   /* 1: joint 0 is a free-flyer (nu = 6) */
    int nframes, ntasks, ncontacts;
    int nl0, nl1;        /* number of level-0 / level-1 blocks */
    int nref;            /* doubles in one instance's reference vector */
    int sc_jacobian_world; /* 1: self-collision uses the spatial WORLD Jacobian, 0: LOCAL_WORLD_ALIGNED */
    double dt;
    double gravity[3];   /* (0, 0, -9.81) */
    double hess_reg;     /* 1e-8 (SolverHQuadProgFast) */
    double total_mass;
    /* ---- tree: joint j in [0, nj); parent[j] < j or -1 (universe) */
    int jtype[IWBC_MAXJ];
    int parent[IWBC_MAXJ];
    int idx_q[IWBC_MAXJ], idx_v[IWBC_MAXJ];
    int depth[IWBC_MAXJ];            /* 0 for children of the universe */
    int last_desc[IWBC_MAXJ];        /* index of the last joint in j's subtree (joints are DFS-ordered) */
    double jplace[IWBC_MAXJ][12];    /* placement in the parent joint frame: R row-major (9), p (3) */
    double axis[IWBC_MAXJ][3];       /* unit axis in the joint frame */
    double mass[IWBC_MAXJ];
    double com[IWBC_MAXJ][3];        /* in the joint frame */
    double inertia[IWBC_MAXJ][6];    /* about the CoM, joint-frame axes: xx xy xz yy yz zz */
    /* ---- limits of the na actuated dofs (tail of v) */
    double q_lo[IWBC_MAXNV], q_hi[IWBC_MAXNV], v_max[IWBC_MAXNV], tau_max[IWBC_MAXNV];
    /* ---- frames used by tasks / contacts */
    int fparent[IWBC_MAXF];          /* supporting joint, -1 = universe */
    double fplace[IWBC_MAXF][12];    /* placement in the supporting joint frame */
    /* ---- task / contact tables and HQP ordering */
    iwbc_task_t tasks[IWBC_MAXT];
    iwbc_contact_t contacts[IWBC_MAXC];
    iwbc_block_t l0[IWBC_MAXL0];     /* level 0 in insertion order (base dynamics first) */
    iwbc_block_t l1[IWBC_MAXL1];     /* level 1 in insertion order = Hessian summation order */
    /* ---- construction-time state */
    double q0[IWBC_MAXNV + 1];       /* reference configuration (pinocchio layout, quaternion normalised) */
    /* ---- force-regularisation weight vectors: [0] = Contact6d's default (1, 1, 1e-3, 2, 2, 2); the others are the stabilizer
     * files' zmp_w (HumanoidPosTracker::update installs them with setRegularizationTaskWeightVector,
     * src/controllers/humanoid_pos_tracker.cpp:220-221).  An instance selects one per contact through its reference row. */
    int nwreg;
    int pad2_;
    double wreg_variants[IWBC_MAXWV][6];
    /* ---- stabilizer (HumanoidPosTracker::parse_stabilizer, src/controllers/humanoid_pos_tracker.cpp:81-114) */
    int stabilizer_activated;        /* CONTROLLER.stabilizer.activated */
    int stab_present;                /* 1: the three parameter sets below were read (humanoid controllers) */
    int stab_talos;                  /* 1: TalosPosTracker::update's IMU momentum admittance applies (talos_pos_tracker.cpp:135-150) */
    int stab_task_com, stab_task_lf, stab_task_rf, stab_task_torso, stab_task_momentum; /* task indices, -1: absent */
    int stab_contact_lf, stab_contact_rf; /* contact indices of contact_lfoot / contact_rfoot, -1: absent */
    int stab_imu_frame;              /* frame table index of imu_link, -1: absent */
    int pad3_;
    iwbc_stab_conf_t stab[3];        /* indexed by iwbc_behavior_type */
} iwbc_problem_t;

#define IWBC_PROBLEM_MAGIC 0x43425749u
#define IWBC_PROBLEM_VERSION 2u

/* layout of the per-instance reference vector (doubles):
 *   SE3 task           : pos[12] = p(3) ++ R column-major(9), vel[6], acc[6]   (24)
 *                        (TrajectorySample(12,6): include/inria_wbc/trajs/utils.hpp:28-36)
 *   contact            : the same SE3 sample (24), then the contact's force reference f_ref[6] (Contact6d::setForceReference,
 *                        0 by default), the index of its regularisation weight vector in wreg_variants (stored as a double,
 *                        0 by default) and one pad                                (32)
 *   COM task           : pos[3], vel[3], acc[3]                                  (9)
 *   MOMENTUM task      : value[6] (unused), derivative[6] = L_ref, second[6] = dL_ref (18)
 *   POSTURE task       : pos[na], vel[na], acc[na]                               (3 na)
 */
#define IWBC_REF_SE3 24
#define IWBC_REF_CONTACT 32
#define IWBC_REF_CONTACT_FREF 24   /* offset of f_ref inside a contact's block */
#define IWBC_REF_CONTACT_WREG 30   /* offset of the weight-vector index */
#define IWBC_REF_COM 9
#define IWBC_REF_MOM 18

/* per-instance status codes = tsid HQP status (src/controllers/controller.cpp:287-304) */
enum iwbc_status {
    IWBC_STATUS_UNKNOWN = -1,
    IWBC_STATUS_OPTIMAL = 0,
    IWBC_STATUS_INFEASIBLE = 1,
    IWBC_STATUS_UNBOUNDED = 2,
    IWBC_STATUS_MAX_ITER = 3,
    IWBC_STATUS_ERROR = 4
};

#ifdef __cplusplus
}
#endif
#endif /* IWBC_PROBLEM_H */
