"""ctypes mirror of include/iwbc_problem.h (read-only view of a compiled problem)."""
import ctypes as C

MAXJ, MAXNV, MAXF, MAXT, MAXC, MAXAV, MAXL0, MAXL1, MAXWV = 56, 64, 48, 32, 4, 12, 16, 40, 4


class Task(C.Structure):
    _fields_ = [("type", C.c_int), ("level", C.c_int), ("nrows", C.c_int), ("frame", C.c_int), ("ref_off", C.c_int), ("nav", C.c_int),
                ("av_frame", C.c_int * MAXAV), ("mask", C.c_ubyte * MAXNV), ("weight", C.c_double), ("kp", C.c_double * 6),
                ("kd", C.c_double * 6), ("av_r0", C.c_double * MAXAV), ("radius", C.c_double), ("sc_k", C.c_double),
                ("sc_sp", C.c_double), ("sc_m", C.c_double)]


class Contact(C.Structure):
    _fields_ = [("frame", C.c_int), ("ref_off", C.c_int), ("active_default", C.c_int), ("pad_", C.c_int), ("kp", C.c_double * 6),
                ("kd", C.c_double * 6), ("T", C.c_double * 72), ("B", C.c_double * 204), ("lb", C.c_double * 17), ("ub", C.c_double * 17),
                ("wreg", C.c_double * 6), ("fref", C.c_double * 6), ("w_forcereg", C.c_double)]


class Block(C.Structure):
    _fields_ = [("kind", C.c_int), ("index", C.c_int)]


class StabConf(C.Structure):
    _fields_ = [("filter_size", C.c_int), ("use_zmp", C.c_int), ("use_momentum", C.c_int), ("zmp_w_variant", C.c_int),
                ("com_gains", C.c_double * 6), ("ankle_gains", C.c_double * 6), ("ffda_gains", C.c_double * 3), ("pad_", C.c_double),
                ("zmp_p", C.c_double * 6), ("zmp_d", C.c_double * 6), ("zmp_w", C.c_double * 6), ("momentum_p", C.c_double * 6),
                ("momentum_d", C.c_double * 6)]


class ProblemStruct(C.Structure):
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
                ("tasks", Task * MAXT), ("contacts", Contact * MAXC), ("l0", Block * MAXL0), ("l1", Block * MAXL1),
                ("q0", C.c_double * (MAXNV + 1)),
                ("nwreg", C.c_int), ("pad2_", C.c_int), ("wreg_variants", (C.c_double * 6) * MAXWV),
                ("stabilizer_activated", C.c_int), ("stab_present", C.c_int), ("stab_talos", C.c_int), ("stab_task_com", C.c_int),
                ("stab_task_lf", C.c_int), ("stab_task_rf", C.c_int), ("stab_task_torso", C.c_int), ("stab_task_momentum", C.c_int),
                ("stab_contact_lf", C.c_int), ("stab_contact_rf", C.c_int), ("stab_imu_frame", C.c_int), ("pad3_", C.c_int),
                ("stab", StabConf * 3)]


def view(ptr):
    return C.cast(ptr, C.POINTER(ProblemStruct)).contents
