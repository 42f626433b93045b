// cpu_sim.cpp -- TEST TOOLING: runs the *device source* of the fused cycle (inria_wbc_b200/csrc/cycle_kernel.cuh)
// on the host with a group of ONE thread (IWBC_CPU_SIM), so that the control flow and arithmetic of
// the CUDA path can be checked against the oracle in the CPU test-suite (-m "not gpu").
// Not part of the product: libiwbc_b200.so never contains this code path.
#define IWBC_CPU_SIM 1
#include <algorithm>
#include <cstdio>
#include <vector>

#include "../inria_wbc_b200/csrc/cycle_kernel.cuh"
#include "../inria_wbc_b200/csrc/fast_kernel.cuh"
#include "../inria_wbc_b200/csrc/host_state.hpp"

namespace iwbc {
void set_error(const std::string&) {}
}
using namespace iwbc;

// same shape preparation as batch.cu::prepare_shape (kept in sync by tests/test_cpu_sim.py comparing results)
static void prepare(const iwbc_problem_t& P, unsigned mask, dev::Shape& S, dev::Sizes& z)
{
    std::memset(&S, 0, sizeof(S));
    host::QPShape qs = host::qp_shape(P, mask);
    S.n = qs.n, S.nk = qs.nk, S.me = qs.me, S.mi = qs.mi;
    int nk = 0;
    for (int c = 0; c < IWBC_MAXC; ++c)
        S.cidx[c] = -1;
    for (int c = 0; c < P.ncontacts; ++c)
        if (mask & (1u << c)) {
            S.cidx[c] = nk;
            nk += 12;
        }
    int ie = 0, ii = 0;
    for (int k = 0; k < P.nl0; ++k) {
        const iwbc_block_t& B = P.l0[k];
        if (B.kind == IWBC_B_BASEDYN) {
            S.ce[S.nce++] = {0, 0, ie, P.nv - P.na};
            ie += P.nv - P.na;
        }
        else if (B.kind == IWBC_B_CONTACT_MOTION) {
            if (S.cidx[B.index] < 0)
                continue;
            S.ce[S.nce++] = {1, B.index, ie, 6};
            ie += 6;
        }
        else if (B.kind == IWBC_B_CONTACT_FORCE) {
            if (S.cidx[B.index] < 0)
                continue;
            S.ci[S.nci++] = {2, B.index, ii, 17};
            ii += 34;
        }
        else {
            const iwbc_task_t& T = P.tasks[B.index];
            S.ci[S.nci++] = {T.type == IWBC_T_BOUNDS ? 0 : 1, B.index, ii, T.nrows};
            ii += 2 * T.nrows;
        }
    }
    for (int k = 0; k < P.nl1; ++k)
        if (P.l1[k].kind == IWBC_B_TASK && P.tasks[P.l1[k].index].type != IWBC_T_POSTURE && P.tasks[P.l1[k].index].type != IWBC_T_TORQUE && P.tasks[P.l1[k].index].type != IWBC_T_COP)
            S.rows_dense += P.tasks[P.l1[k].index].nrows;
    for (int j = 0; j < P.nj; ++j)
        S.maxdepth = std::max(S.maxdepth, P.depth[j]);
    S.ldJ = S.n | 1;
    S.ldM = P.nv | 1;
    z.nv = P.nv, z.nq = P.nq, z.nj = P.nj, z.nf = P.nframes, z.nc = P.ncontacts, z.n = S.n, z.mi = S.mi;
    z.rows_dense = S.rows_dense, z.nref = P.nref, z.ldJ = S.ldJ, z.ldM = S.ldM;
    dev::make_layout(z, S);
}

extern "C" int sim_cycle_batch(const iwbc_problem_t* P, int B, const double* q, const double* v, const double* refs, long refs_stride,
    unsigned mask, double* tau, double* dv, double* f, double* q_next, double* v_next, int* status, int* iters, unsigned* active,
    double* q_cmd, double* momentum)
{
    dev::Sizes z;
    static dev::Shape S;
    prepare(*P, mask, S, z);
    z.S = &S, z.P = P;
    std::vector<double> smem(S.total + 16, 0.0);
    dev::Sm sm = dev::make_sm(smem.data(), z);
    for (int j = 0; j < z.nj; ++j) {
        const int c0 = P->idx_v[j], cn = P->jtype[j] == IWBC_J_FREEFLYER ? 6 : 1;
        for (int k = 0; k < cn; ++k)
            sm.dofj[c0 + k] = j;
    }
    dev::IO io;
    std::memset(&io, 0, sizeof(io));
    io.B = B, io.q = q, io.v = v, io.refs = refs, io.refs_stride = refs_stride;
    io.tau = tau, io.dv = dv, io.f = f, io.q_next = q_next, io.v_next = v_next, io.status = status, io.iters = iters, io.active = active;
    io.q_cmd = q_cmd, io.momentum = momentum;
    int flip = 0;
    for (int i = 0; i < B; ++i)
        dev::cycle_instance<1>(P, sm, z, io, i, 0, flip);
    return 0;
}

// fast path (fast_kernel.cuh) with the active-set kernel as continuation, exactly as batch.cu chains them.
// fast_used (optional, [B]): 1 where the instance was finished by the fast path.
extern "C" int sim_fast_batch(const iwbc_problem_t* P, int B, const double* q, const double* v, const double* refs, long refs_stride,
    unsigned mask, double* tau, double* dv, double* f, double* q_next, double* v_next, int* status, int* iters, unsigned* active,
    double* q_cmd, double* momentum, int* fast_used)
{
    dev::Sizes z;
    static dev::Shape S;
    static dev::FShape F;
    prepare(*P, mask, S, z);
    z.S = &S, z.P = P;
    static dev::FPlan plan;
    if (!dev::build_fast_plan(*P, S, 1, F, plan))
        return -1;
    std::vector<double> smem(S.total + 16, 0.0), fsmem(F.s_total + 16, 0.0);
    dev::Sm sm = dev::make_sm(smem.data(), z);
    int* fdofj = reinterpret_cast<int*>(fsmem.data() + F.o_int);
    for (int j = 0; j < z.nj; ++j) {
        const int c0 = P->idx_v[j], cn = P->jtype[j] == IWBC_J_FREEFLYER ? 6 : 1;
        for (int k = 0; k < cn; ++k)
            sm.dofj[c0 + k] = j, fdofj[c0 + k] = j;
    }
    dev::IO io;
    std::memset(&io, 0, sizeof(io));
    io.B = B, io.q = q, io.v = v, io.refs = refs, io.refs_stride = refs_stride;
    io.tau = tau, io.dv = dv, io.f = f, io.q_next = q_next, io.v_next = v_next, io.status = status, io.iters = iters, io.active = active;
    io.q_cmd = q_cmd, io.momentum = momentum;
    int flip = 0;
    for (int i = 0; i < B; ++i) {
        const int code = dev::fast_instance<1>(P, &plan, S, F, fsmem.data(), io, i, 0, nullptr);
        if (fast_used)
            fast_used[i] = code > 0 ? 1 : 0;
        if (code <= 0)
            dev::slow_continue<1>(P, &plan, S, F, fsmem.data(), io, i, 0, code, flip, nullptr);
    }
    return 0;
}

extern "C" int sim_qp_solve(int n, int me, int mi, const double* H, const double* g, const double* CE, const double* ce0, const double* CI,
    const double* ci0, double* x, int* status, int* iters, unsigned* active)
{
    static dev::Shape S;
    std::memset(&S, 0, sizeof(S));
    S.n = n, S.me = me, S.mi = mi, S.ldJ = n | 1, S.ldM = 1;
    dev::Sizes z;
    std::memset(&z, 0, sizeof(z));
    z.n = n, z.mi = mi, z.ldJ = S.ldJ, z.ldM = 1;
    z.S = &S;
    dev::make_layout(z, S);
    std::vector<double> smem(S.total + 16, 0.0);
    dev::Sm sm = dev::make_sm(smem.data(), z);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j)
            sm.J[i + j * z.ldJ] = H[i * n + j];
    for (int k = 0; k < n; ++k)
        sm.g[k] = g[k];
    dev::DenseRows rows = {CE, ce0, CI, ci0, n, me, mi};
    int flip = 0, it = 0, iq = 0;
    *status = dev::gi_solve<1, dev::DenseRows>(nullptr, sm, z, &rows, me, mi, 0, flip, it, iq);
    for (int k = 0; k < n; ++k)
        x[k] = sm.x[k];
    if (iters)
        *iters = it;
    if (active) {
        for (int k = 0; k < IWBC_ACTIVE_WORDS; ++k)
            active[k] = 0;
        for (int k = me; k < iq; ++k) {
            int r = sm.act[k];
            if (r >= 0 && r < 32 * IWBC_ACTIVE_WORDS)
                active[r >> 5] |= 1u << (r & 31);
        }
    }
    return 0;
}

// shared-memory footprints (bytes) of the fast pass / active-set continuation / generic kernel for a contact mask (T = 128 plan)
extern "C" int sim_smem_bytes(const iwbc_problem_t* P, unsigned mask, long* out3)
{
    dev::Sizes z;
    static dev::Shape S;
    static dev::FShape F;
    static dev::FPlan plan;
    prepare(*P, mask, S, z);
    out3[2] = (long)S.total * 8;
    out3[0] = out3[1] = 0;
    if (!dev::build_fast_plan(*P, S, 128, F, plan))
        return -1;
    out3[0] = (long)F.total * 8, out3[1] = (long)F.s_total * 8;
    return 0;
}
