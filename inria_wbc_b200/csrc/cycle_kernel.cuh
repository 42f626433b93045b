// cycle_kernel.cuh -- the fused batched control cycle for sm_100a.
//
// One *group* of G threads steps one robot state through the whole inria_wbc update
// (/root/reference/src/controllers/controller.cpp:231-313):
//   k_dyn   world-frame kinematics / dynamics of the URDF tree: FK, joint Jacobian, CRBA mass matrix,
//           RNEA non-linear effects, centroidal map, drift terms           (pinocchio computeAllTerms)
//   k_asm   task + contact rows, H = sum w A'A + 1e-8 I, g = -sum w A'b    (TSID formulation + SolverHQuadProgFast)
//   k_qp    Goldfarb-Idnani dual active set with J, R resident in shared memory (EiquadprogFast)
//   k_dec   tau = h_a + M_a dv - J_a' f, v+ = v + dt dv, q+ = q (+) dt v+, command repack
// all inside ONE kernel so that H (n^2 doubles) never round-trips through HBM; per instance the
// kernel reads (q, v, refs) and writes (tau, dv, f, q+, v+, status): a few kB.
//
// Mapping: G = 96 threads = 1 CTA per instance for humanoids (n = 62..74: one thread per row /
// column of J), G = 32 (one warp) or less per instance, several instances per CTA, for arms
// (n <= 16).  Groups are persistent: they stride over the batch.  No atomics, fixed summation
// orders => bitwise run-to-run determinism (the reference's test_determinism contract).
//
// Numerics follow the reference algorithms but not their data layout:
//   * dynamics are evaluated in world coordinates (spatial vectors about the world origin), so
//     subtree quantities are plain sums over the DFS-contiguous joint range of the subtree;
//   * constraint rows are never materialised: base-dynamics / contact / bound / actuation-bound /
//     friction rows are evaluated from M, the contact Jacobians and constant 17x12 blocks;
//   * the Givens sweep of add_constraint is applied row-parallel: rotation coefficients come from
//     suffix sums of d^2 (same association order as the sequential hypot chain), then each thread
//     rotates its own row of J;
//   * multipliers of *equality* constraints are not tracked (they never feed back into x, the
//     active set or the status).
#pragma once
#ifdef IWBC_CPU_SIM
// Debug build of this very source for a host compiler: a "group" of ONE thread executes the kernel
// body serially (tests/cpu_sim.cpp).  It exercises the control flow and arithmetic of the device
// code without a GPU; it is test tooling, not a product path (the shipped library never defines
// IWBC_CPU_SIM).
#include <cmath>
#include <cstring>
#include <type_traits>
#define __device__
#define __host__
#define __noinline__
#define __forceinline__ inline
#define __constant__
#define __grid_constant__
#define __restrict__
namespace iwbc { namespace dev {
struct SimIdx { int x; };
static SimIdx threadIdx = {0};
inline void __syncwarp(unsigned) {}
inline void __syncthreads() {}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> inline T __shfl_down_sync(unsigned, T v, int, int = 32) { return v; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const unsigned long long o = *p; *p += v; return o; }
inline int atomicAdd(int* p, int v) { const int o = *p; *p += v; return o; }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline long long clock64() { return 0; }
inline void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
using std::sqrt; using std::fabs; using std::fmin; using std::fmax; using std::exp; using std::log1p; using std::acos; using std::sin; using std::atan2; using std::fma;
} }
#else
#include <cuda_runtime.h>
#endif

#include "../../include/iwbc_batch.h"

namespace iwbc {
namespace dev {

    // ------------------------------------------------------------------ launch-constant data
    struct CiBlock {
        int kind;  // 0 bounds, 1 actuation bounds, 2 friction (contact `index`)
        int index; // task or contact index
        int off;   // first CI row
        int rows;  // rows of ONE side (the block has 2*rows rows: +A then -A)
    };
    struct CeBlock {
        int kind; // 0 base dynamics, 1 contact motion
        int index;
        int off;
        int rows;
    };
    struct Shape {
        int n, nk, me, mi, rows_dense; // QP sizes for this launch's contact mask; dense level-1 rows (everything but posture / force-reg)
        int nce, nci;
        int cidx[IWBC_MAXC]; // force-column offset of contact c (-1: inactive)
        int maxdepth;
        int ldJ, ldM;
        CeBlock ce[IWBC_MAXL0];
        CiBlock ci[IWBC_MAXL0];
        // shared-memory layout (offsets in doubles, per instance)
        int o_J, o_R, o_M, o_Jft, o_cb, o_nle, o_vec, o_ovl, o_dyn, o_A, o_b, o_int, total;
    };

    struct IO {
        int B;
        const double* q;
        const double* v;
        const double* refs;
        long refs_stride;
        double* tau;
        double* dv;
        double* f;
        double* q_next;
        double* v_next;
        double* q_cmd;
        double* momentum;
        int* status;
        int* iters;
        unsigned* active;
        // rollout: write q_next / v_next back into q / v (in place) when set
        double* q_rw;
        double* v_rw;
        int sticky_status; // roll-outs, ticks after the first: a non-OPTIMAL status already stored for the instance is kept
        // debug dump of instance 0 (row-major): M nle H g x
        double* dbg_M;
        double* dbg_nle;
        double* dbg_H;
        double* dbg_g;
        double* dbg_x;
        double* dbg_A;         // optional: the dense level-1 rows [rows_dense][nv] (scaled by sqrt(weight)) and their right-hand sides
        double* dbg_b;
        long long* dbg_cycles; // phase timestamps (clock64) of instance 0: load, dyn, asm, factor, eq, ineq, decode
        const int* list;       // optional: instance indices to process (active-set continuation of the fast path)
        const int* list_count; // device-side length of `list`
        int* work;             // optional: device counter (zeroed before the launch) for dynamic hand-out of list / batch positions to the groups
        const int* in_list;    // optional: the instances of this launch set (a behaviour's per-contact-mask bucket), device-side length
        const int* in_count;
        int* slow_list;        // fast path: instances that need the active-set iteration are appended here
        int* slow_count;
        double* scratch;       // fast path -> continuation: shared-memory image of the front end of the first scratch_cap listed instances
        int* slow_code;        // ... and their fast_instance() return code
        int scratch_cap;
        unsigned long long* prof; // steady-state phase profile: [0..5] summed SM cycles of dyn, asm, factor, eq, ineq, decode; [6] instances
    };


    // ------------------------------------------------------------------ small device math
    struct V3 {
        double x, y, z;
    };
    __device__ __forceinline__ V3 mk(double x, double y, double z) { return V3{x, y, z}; }
    __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
    __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
    __device__ __forceinline__ V3 operator*(double s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
    __device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
    __device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
    __device__ __forceinline__ V3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }
    __device__ __forceinline__ void st3(double* p, V3 a)
    {
        p[0] = a.x;
        p[1] = a.y;
        p[2] = a.z;
    }
    // R row-major 9
    __device__ __forceinline__ V3 rmul(const double* R, V3 v)
    {
        return mk(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z);
    }
    __device__ __forceinline__ V3 rtmul(const double* R, V3 v)
    {
        return mk(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z);
    }
    __device__ __forceinline__ void mm3(const double* A, const double* B, double* C)
    {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
                C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    }
    __device__ __forceinline__ void axis_angle_R(V3 a, double th, double* R)
    {
        double s, c;
        sincos(th, &s, &c);
        double v = 1 - c;
        R[0] = a.x * a.x * v + c;
        R[1] = a.x * a.y * v - a.z * s;
        R[2] = a.x * a.z * v + a.y * s;
        R[3] = a.x * a.y * v + a.z * s;
        R[4] = a.y * a.y * v + c;
        R[5] = a.y * a.z * v - a.x * s;
        R[6] = a.x * a.z * v - a.y * s;
        R[7] = a.y * a.z * v + a.x * s;
        R[8] = a.z * a.z * v + c;
    }
    __device__ __forceinline__ void quat_R(double x, double y, double z, double w, double* R)
    {
        R[0] = 1 - 2 * (y * y + z * z);
        R[1] = 2 * (x * y - z * w);
        R[2] = 2 * (x * z + y * w);
        R[3] = 2 * (x * y + z * w);
        R[4] = 1 - 2 * (x * x + z * z);
        R[5] = 2 * (y * z - x * w);
        R[6] = 2 * (x * z - y * w);
        R[7] = 2 * (y * z + x * w);
        R[8] = 1 - 2 * (x * x + y * y);
    }
    // log3 / log6 (pinocchio conventions; see oracle/wbc_oracle.cpp for the CPU restatement)
    __device__ inline V3 log3(const double* R)
    {
        double tr = R[0] + R[4] + R[8];
        double c = fmin(1.0, fmax(-1.0, 0.5 * (tr - 1)));
        double theta = acos(c);
        V3 v = mk(R[7] - R[5], R[2] - R[6], R[3] - R[1]);
        if (theta < 1.220703125e-4) // TaylorSeriesExpansion<double>::precision<3>() = eps^(1/4)
            return 0.5 * v;
        if (theta > M_PI - 1e-2) {
            double cphi = -(tr - 1) / 2;
            double beta = theta * theta / (1 + cphi);
            double tx = (R[0] + cphi) * beta, ty = (R[4] + cphi) * beta, tz = (R[8] + cphi) * beta;
            return mk((v.x > 0 ? 1 : -1) * (tx > 0 ? sqrt(tx) : 0), (v.y > 0 ? 1 : -1) * (ty > 0 ? sqrt(ty) : 0),
                (v.z > 0 ? 1 : -1) * (tz > 0 ? sqrt(tz) : 0));
        }
        return (theta / (2 * sin(theta))) * v;
    }
    __device__ inline void log6(const double* R, V3 p, V3& lin, V3& ang)
    {
        V3 w = log3(R);
        double t2 = dot(w, w), t = sqrt(t2);
        double alpha, beta;
        if (t < 1.220703125e-4) {
            alpha = 1 - t2 / 12 - t2 * t2 / 720;
            beta = 1.0 / 12 + t2 / 720;
        }
        else {
            double st, ct;
            sincos(t, &st, &ct);
            alpha = t * st / (2 * (1 - ct));
            beta = 1 / t2 - st / (2 * t * (1 - ct));
        }
        lin = alpha * p - 0.5 * cross(w, p) + (beta * dot(w, p)) * w;
        ang = w;
    }

    // ------------------------------------------------------------------ group primitives
    template <int G>
    struct Grp {
        // G <= 32: a sub-warp (or one warp) per instance, several per CTA; G > 32: the CTA is the group
        static __device__ __forceinline__ unsigned mask()
        {
            if (G >= 32)
                return 0xffffffffu;
            unsigned lane = threadIdx.x & 31;
            return ((G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u))) << (lane / G * G);
        }
        static __device__ __forceinline__ void sync()
        {
            if (G <= 32)
                __syncwarp(mask());
            else
                __syncthreads();
        }
    };

    // sum K values over the group; result broadcast to every thread.  red: 2 * 4 * K doubles of
    // scratch (only for G > 32); `flip` alternates the two halves so one barrier per call suffices.
    template <int G, int K>
    __device__ __forceinline__ void gsum(double (&v)[K], double* red, int& flip)
    {
        const unsigned m = Grp<G>::mask();
        if (G <= 32) {
#pragma unroll
            for (int off = (G < 32 ? G : 32) / 2; off > 0; off >>= 1)
#pragma unroll
                for (int k = 0; k < K; ++k)
                    v[k] += __shfl_xor_sync(m, v[k], off, G < 32 ? G : 32);
        }
        else {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int k = 0; k < K; ++k)
                    v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
            double* r = red + flip * 16;
            flip ^= 1;
            const int w = threadIdx.x >> 5;
            if ((threadIdx.x & 31) == 0)
#pragma unroll
                for (int k = 0; k < K; ++k)
                    r[w * K + k] = v[k];
            __syncthreads();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double s = r[k];
                for (int ww = 1; ww < G / 32; ++ww)
                    s += r[ww * K + k];
                v[k] = s;
            }
        }
    }

    // (value, index) arg-min with "first index wins" ties, broadcast
    template <int G>
    __device__ __forceinline__ void gargmin(double& val, int& idx, double* red, int& flip)
    {
        const unsigned m = Grp<G>::mask();
        const int W = G < 32 ? G : 32;
#pragma unroll
        for (int off = W / 2; off > 0; off >>= 1) {
            double ov = __shfl_xor_sync(m, val, off, W);
            int oi = __shfl_xor_sync(m, idx, off, W);
            if (ov < val || (ov == val && oi < idx)) {
                val = ov;
                idx = oi;
            }
        }
        if (G > 32) {
            double* r = red + flip * 16;
            flip ^= 1;
            const int w = threadIdx.x >> 5;
            if ((threadIdx.x & 31) == 0) {
                r[2 * w] = val;
                r[2 * w + 1] = (double)idx;
            }
            __syncthreads();
            val = r[0];
            idx = (int)r[1];
            for (int ww = 1; ww < G / 32; ++ww) {
                double ov = r[2 * ww];
                int oi = (int)r[2 * ww + 1];
                if (ov < val || (ov == val && oi < idx)) {
                    val = ov;
                    idx = oi;
                }
            }
        }
    }

    // ------------------------------------------------------------------ shared-memory view
    struct Sm {
        double* J;    // n x ldJ column-major: H (lower) -> L -> J = L^-T
        double* R;    // packed upper triangular, column-major: R(i,j) at j(j+1)/2 + i
        double* M;    // nv x ldM (symmetric, both halves); the fast pass / continuation keep one half only: mass_at()
        double* Md;   // diagonal of M when Sizes::msym == 1
        double* Jft;  // [ncontacts][6][nv] local contact-frame Jacobians
        double* cb;   // [ncontacts][6] contact motion right-hand sides; then wrench scratch [ncontacts][6]
        double* nle;  // nv
        // QP / misc vectors (each n+2 long unless noted)
        double *g, *x, *d, *z, *np, *r, *u, *x_old, *u_old, *gc, *gs, *gx, *s, *lb, *ub, *Ldiag, *sub, *red, *sc;
        // dynamics scratch (aliases the QP vectors' tail + R): see layout()
        double *oMi, *Sw, *vw, *aw, *Yw, *fw, *Fc, *oMf, *vf, *af, *qs, *vs, *ref;
        double *A, *b; // dense level-1 rows (alias of R)
        int *act, *act_old, *iai, *dofj, *ints;
    };

    __host__ __device__ inline int align2(int x) { return (x + 1) & ~1; }

    // The layout is computed on the host (fills Shape::o_* and total) and replayed on the device.
    struct Sizes {
        int nv, nq, nj, nf, nc, n, mi, rows_dense, nref, ldJ, ldM;
        int msym = 0; // storage of the mass matrix: see mass_at()
        // per-launch tables: the QP shape of this launch's contact mask (a __grid_constant__ kernel parameter) and the problem
        // (the handle's own global-memory copy, read through the read-only path).  No process-global __constant__ state: two
        // handles, or two contact masks of one handle, never share anything on the device.
        const Shape* S;
        const iwbc_problem_t* P;
    };
    __host__ __device__ inline int vec_len(const Sizes& z) { return align2(z.n + 2); }
    // QP vectors: g x d z np r u x_old u_old gc gs gx (12) + Ldiag sub (2) -> 14 * vec_len ; s (mi) lb ub (na<=nv each)
    __host__ __device__ inline int qp_vec_doubles(const Sizes& z) { return 14 * vec_len(z) + align2(z.mi) + 2 * align2(z.nv) + 32 /*red*/ + 96 /*sc*/; }
    __host__ __device__ inline int dyn_doubles(const Sizes& z)
    {
        return 12 * z.nj + 6 * z.nv + 6 * z.nj + 6 * z.nj + 10 * z.nj + 6 * z.nj + 6 * z.nv + 12 * z.nf + 6 * z.nf + 6 * z.nf + align2(z.nq) + align2(z.nv)
            + align2(z.nref);
    }
    __host__ __device__ inline int int_count(const Sizes& z) { return 2 * (z.n + 2) + z.mi + z.nv + 16; }

    inline void make_layout(const Sizes& z, Shape& S)
    {
        int o = 0;
        S.o_J = o;
        o += align2(z.n * z.ldJ);
        S.o_R = o;
        int rdoubles = z.n * (z.n + 1) / 2;
        int adoubles = z.rows_dense * z.nv + align2(z.rows_dense);
        o += align2(rdoubles > adoubles ? rdoubles : adoubles);
        S.o_A = S.o_R;
        S.o_b = S.o_R + z.rows_dense * z.nv;
        S.o_M = o;
        o += align2(z.nv * z.ldM);
        S.o_Jft = o;
        o += align2(z.nc * 6 * z.nv);
        S.o_cb = o;
        o += align2(z.nc * 12);
        S.o_nle = o;
        o += align2(z.nv);
        // vectors alive from the assembly on: g, lb, ub, red, sc
        S.o_vec = o;
        o += vec_len(z) + 2 * align2(z.nv) + 32 + 96;
        // overlay: the 13 vectors only the QP uses (x d z np r u x_old u_old gc gs gx Ldiag sub) and s share their bytes with the
        // dynamics scratch that dies with the assembly (everything of dyn_doubles() but qs, vs): the warp-per-instance kernel is
        // latency bound and its occupancy is set by shared memory per instance (Tiago: 60 kB -> 54 kB per CTA = 4 CTAs per SM)
        S.o_ovl = o;
        {
            const int qp_only = 13 * vec_len(z) + align2(z.mi);
            const int dyn_dead = dyn_doubles(z) - align2(z.nq) - align2(z.nv);
            o += qp_only > dyn_dead ? qp_only : dyn_dead;
        }
        S.o_dyn = o; // qs, vs
        o += align2(z.nq) + align2(z.nv);
        S.o_int = o;
        o += align2((int_count(z) + 1) / 2);
        S.total = o;
    }

    __device__ __forceinline__ Sm make_sm(double* base, const Sizes& z)
    {
        const Shape& S = *z.S;
        Sm s;
        s.J = base + S.o_J;
        s.R = base + S.o_R;
        s.A = base + S.o_A;
        s.b = base + S.o_b;
        s.M = base + S.o_M;
        s.Jft = base + S.o_Jft;
        s.cb = base + S.o_cb;
        s.nle = base + S.o_nle;
        double* p = base + S.o_vec;
        const int L = vec_len(z);
        s.g = p, p += L;
        s.lb = p, p += align2(z.nv);
        s.ub = p, p += align2(z.nv);
        s.red = p, p += 32;
        s.sc = p, p += 96;
        p = base + S.o_ovl; // QP-only vectors ...
        s.x = p, p += L;
        s.d = p, p += L;
        s.z = p, p += L;
        s.np = p, p += L;
        s.r = p, p += L;
        s.u = p, p += L;
        s.x_old = p, p += L;
        s.u_old = p, p += L;
        s.gc = p, p += L;
        s.gs = p, p += L;
        s.gx = p, p += L;
        s.Ldiag = p, p += L;
        s.sub = p, p += L;
        s.s = p, p += align2(z.mi);
        p = base + S.o_ovl; // ... over the dynamics scratch that is dead once H, g and the constraint data are assembled
        s.oMi = p, p += 12 * z.nj;
        s.Sw = p, p += 6 * z.nv;
        s.vw = p, p += 6 * z.nj;
        s.aw = p, p += 6 * z.nj;
        s.Yw = p, p += 10 * z.nj;
        s.fw = p, p += 6 * z.nj;
        s.Fc = p, p += 6 * z.nv;
        s.oMf = p, p += 12 * z.nf;
        s.vf = p, p += 6 * z.nf;
        s.af = p, p += 6 * z.nf;
        s.ref = p, p += align2(z.nref);
        p = base + S.o_dyn;
        s.qs = p, p += align2(z.nq);
        s.vs = p, p += align2(z.nv);
        int* ip = reinterpret_cast<int*>(base + S.o_int);
        s.act = ip, ip += z.n + 2;
        s.act_old = ip, ip += z.n + 2;
        s.iai = ip, ip += z.mi;
        s.dofj = ip, ip += z.nv;
        s.ints = ip;
        return s;
    }

    __device__ __forceinline__ int Ridx(int i, int j) { return j * (j + 1) / 2 + i; }

    // does joint a support (is an ancestor-or-self of) joint j?  joints are DFS-numbered
    __device__ __forceinline__ bool supports(const iwbc_problem_t* gP, int a, int j) { return a <= j && j <= gP->last_desc[a]; }

    // =================================================================================== k_dyn
    template <int G>
    __device__ __forceinline__ void k_dyn(const iwbc_problem_t* __restrict__ gP, const Sm& sm, const Sizes& z, int tid, int& flip)
    {
        const int nj = z.nj, nv = z.nv;
        // ---- placements, world motion subspaces, velocities and drift accelerations, level by level
        // (a joint is handled by thread j % G when its depth comes up; parents are one level above)
        for (int dpt = 0; dpt <= z.S->maxdepth; ++dpt) {
            for (int j0 = tid; j0 < nj; j0 += G) {
                if (gP->depth[j0] != dpt)
                    continue;
                double li[12]; // liMi: placement in the parent joint frame times the joint transform
                {
                    const double* jp = gP->jplace[j0];
                    const int ty = gP->jtype[j0];
                    double jm[9];
                    V3 jt = mk(0, 0, 0);
                    if (ty == IWBC_J_FREEFLYER) {
                        const double* qq = sm.qs + gP->idx_q[j0];
                        quat_R(qq[3], qq[4], qq[5], qq[6], jm);
                        jt = mk(qq[0], qq[1], qq[2]);
                    }
                    else if (ty == IWBC_J_REVOLUTE) {
                        axis_angle_R(ld3(gP->axis[j0]), sm.qs[gP->idx_q[j0]], jm);
                    }
                    else {
                        jm[0] = jm[4] = jm[8] = 1;
                        jm[1] = jm[2] = jm[3] = jm[5] = jm[6] = jm[7] = 0;
                        jt = sm.qs[gP->idx_q[j0]] * ld3(gP->axis[j0]);
                    }
                    mm3(jp, jm, li);
                    st3(li + 9, rmul(jp, jt) + ld3(jp + 9));
                }
                const int my_parent = gP->parent[j0];
                double* o = sm.oMi + 12 * j0;
                V3 vp = mk(0, 0, 0), wp = mk(0, 0, 0), alp = mk(0, 0, 0), aap = mk(0, 0, 0);
                if (my_parent >= 0) {
                    const double* pm = sm.oMi + 12 * my_parent;
                    mm3(pm, li, o);
                    st3(o + 9, rmul(pm, ld3(li + 9)) + ld3(pm + 9));
                    vp = ld3(sm.vw + 6 * my_parent);
                    wp = ld3(sm.vw + 6 * my_parent + 3);
                    alp = ld3(sm.aw + 6 * my_parent);
                    aap = ld3(sm.aw + 6 * my_parent + 3);
                }
                else {
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        o[k] = li[k];
                }
                // world motion-subspace columns S = oMi.act(S_local) and joint velocity vJ = S qd
                const V3 p = ld3(o + 9);
                const int ty = gP->jtype[j0], c0 = gP->idx_v[j0];
                V3 vJl = mk(0, 0, 0), vJa = mk(0, 0, 0);
                if (ty == IWBC_J_FREEFLYER) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        V3 col = mk(o[k], o[3 + k], o[6 + k]); // R[:,k]
                        st3(sm.Sw + 6 * (c0 + k), col);
                        st3(sm.Sw + 6 * (c0 + k) + 3, mk(0, 0, 0));
                        V3 lin = cross(p, col);
                        st3(sm.Sw + 6 * (c0 + 3 + k), lin);
                        st3(sm.Sw + 6 * (c0 + 3 + k) + 3, col);
                        vJl = vJl + sm.vs[c0 + k] * col + sm.vs[c0 + 3 + k] * lin;
                        vJa = vJa + sm.vs[c0 + 3 + k] * col;
                    }
                }
                else {
                    V3 aw = rmul(o, ld3(gP->axis[j0]));
                    V3 lin, ang;
                    if (ty == IWBC_J_REVOLUTE) {
                        ang = aw;
                        lin = cross(p, aw);
                    }
                    else {
                        ang = mk(0, 0, 0);
                        lin = aw;
                    }
                    st3(sm.Sw + 6 * c0, lin);
                    st3(sm.Sw + 6 * c0 + 3, ang);
                    vJl = sm.vs[c0] * lin;
                    vJa = sm.vs[c0] * ang;
                }
                // v = v_parent + vJ ; drift acceleration a = a_parent + v_parent x vJ (spatial cross, world coords)
                st3(sm.vw + 6 * j0, vp + vJl);
                st3(sm.vw + 6 * j0 + 3, wp + vJa);
                st3(sm.aw + 6 * j0, alp + cross(vp, vJa) + cross(wp, vJl));
                st3(sm.aw + 6 * j0 + 3, aap + cross(wp, vJa));
            }
            Grp<G>::sync();
        }
        // ---- per joint: world inertia about the origin (m, h = m c, I_O), momentum, force
        for (int j = tid; j < nj; j += G) {
            const double* o = sm.oMi + 12 * j;
            const double m = gP->mass[j];
            V3 c = rmul(o, ld3(gP->com[j])) + ld3(o + 9);
            const double* I = gP->inertia[j];
            double Il[9] = {I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]};
            double t[9], Iw[9];
            mm3(o, Il, t);
            // Iw = t * R^T
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int bb = 0; bb < 3; ++bb)
                    Iw[3 * a + bb] = t[3 * a] * o[3 * bb] + t[3 * a + 1] * o[3 * bb + 1] + t[3 * a + 2] * o[3 * bb + 2];
            const double cc = dot(c, c);
            double Ixx = Iw[0] + m * (cc - c.x * c.x), Ixy = Iw[1] - m * c.x * c.y, Ixz = Iw[2] - m * c.x * c.z;
            double Iyy = Iw[4] + m * (cc - c.y * c.y), Iyz = Iw[5] - m * c.y * c.z, Izz = Iw[8] + m * (cc - c.z * c.z);
            V3 h = m * c;
            double* Y = sm.Yw + 10 * j;
            Y[0] = m;
            st3(Y + 1, h);
            Y[4] = Ixx, Y[5] = Ixy, Y[6] = Ixz, Y[7] = Iyy, Y[8] = Iyz, Y[9] = Izz;
            // momentum Y v and force f = Y a_gf + v x* (Y v), a_gf = a_drift + (-g, 0)
            V3 vl = ld3(sm.vw + 6 * j), w = ld3(sm.vw + 6 * j + 3);
            V3 al = ld3(sm.aw + 6 * j) - ld3(z.P->gravity), aa = ld3(sm.aw + 6 * j + 3);
            auto Iapply = [&](V3 x) { return mk(Ixx * x.x + Ixy * x.y + Ixz * x.z, Ixy * x.x + Iyy * x.y + Iyz * x.z, Ixz * x.x + Iyz * x.y + Izz * x.z); };
            V3 hl = m * vl - cross(h, w), ha = Iapply(w) + cross(h, vl);
            V3 fl = m * al - cross(h, aa) + cross(w, hl);
            V3 fa = Iapply(aa) + cross(h, al) + cross(w, ha) + cross(vl, hl);
            st3(sm.fw + 6 * j, fl);
            st3(sm.fw + 6 * j + 3, fa);
            // body momentum is only needed as a total: keep it in Fc scratch rows (nv >= nj)
            st3(sm.Fc + 6 * j, hl);
            st3(sm.Fc + 6 * j + 3, ha);
        }
        Grp<G>::sync();
        // ---- totals (fixed order): mass/h/I (10), momentum (6), force (6) -> sc[0..21]
        for (int k = tid; k < 22; k += G) {
            double s = 0;
            for (int j = 0; j < nj; ++j)
                s += k < 10 ? sm.Yw[10 * j + k] : k < 16 ? sm.Fc[6 * j + (k - 10)] : sm.fw[6 * j + (k - 16)];
            sm.sc[k] = s;
        }
        Grp<G>::sync();
        // ---- subtree sums (composite inertia, subtree force): joint j sums its DFS range.  In place, so the sums
        // are staged in registers (chunks of 8 items per thread) and written back after a barrier.
        {
            const int items = nj * 16; // (joint, quantity), 16 quantities per joint
            for (int base = 0; base < items; base += 8 * G) {
                double acc[8];
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const int it = base + c8 * G + tid;
                    double s = 0;
                    if (it < items) {
                        const int j = it >> 4, k = it & 15;
                        const int last = gP->last_desc[j];
                        for (int jj = j; jj <= last; ++jj)
                            s += k < 10 ? sm.Yw[10 * jj + k] : sm.fw[6 * jj + (k - 10)];
                    }
                    acc[c8] = s;
                }
                // a later chunk still has to read the un-summed values of joints >= its own, and the DFS ranges only
                // extend forward: chunks are processed in increasing joint order, so writing back chunk c is safe
                Grp<G>::sync();
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const int it = base + c8 * G + tid;
                    if (it < items) {
                        const int j = it >> 4, k = it & 15;
                        if (k < 10)
                            sm.Yw[10 * j + k] = acc[c8];
                        else
                            sm.fw[6 * j + (k - 10)] = acc[c8];
                    }
                }
                Grp<G>::sync();
            }
        }
        // ---- per dof: nle = S . f_subtree, F = Yc S
        for (int c = tid; c < nv; c += G) {
            const int j = sm.dofj[c];
            V3 sl = ld3(sm.Sw + 6 * c), sa = ld3(sm.Sw + 6 * c + 3);
            sm.nle[c] = dot(sl, ld3(sm.fw + 6 * j)) + dot(sa, ld3(sm.fw + 6 * j + 3));
            const double* Y = sm.Yw + 10 * j;
            V3 h = ld3(Y + 1);
            V3 Fl = Y[0] * sl - cross(h, sa);
            V3 Fa = mk(Y[4] * sa.x + Y[5] * sa.y + Y[6] * sa.z, Y[5] * sa.x + Y[7] * sa.y + Y[8] * sa.z, Y[6] * sa.x + Y[8] * sa.y + Y[9] * sa.z) + cross(h, sl);
            st3(sm.Fc + 6 * c, Fl);
            st3(sm.Fc + 6 * c + 3, Fa);
        }
        Grp<G>::sync();
        // ---- mass matrix: M[r][c] = S_lo . F_hi when joint(lo) supports joint(hi), else 0
        for (int e = tid; e < nv * nv; e += G) {
            const int r = e / nv, c = e - r * nv;
            const int lo = r < c ? r : c, hi = r < c ? c : r;
            double m = 0;
            if (supports(gP, sm.dofj[lo], sm.dofj[hi]))
                m = dot(ld3(sm.Sw + 6 * lo), ld3(sm.Fc + 6 * hi)) + dot(ld3(sm.Sw + 6 * lo + 3), ld3(sm.Fc + 6 * hi + 3));
            sm.M[r * z.ldM + c] = m;
        }
        // ---- centroidal quantities: sc[22..24] com, [25..27] vcom, [28..30] acom drift, [31..36] hg, [37..42] dhg drift
        if (tid == 0) {
            const double mt = sm.sc[0];
            V3 com = (1.0 / mt) * ld3(sm.sc + 1);
            V3 hl = ld3(sm.sc + 10), ha = ld3(sm.sc + 13);
            V3 fl = ld3(sm.sc + 16), fa = ld3(sm.sc + 19);
            // drift force = total force - Ytot * a_g, a_g = (-g, 0): lin = m(-g), ang = h x (-g)
            V3 ag = -1.0 * ld3(z.P->gravity);
            V3 dl = fl - mt * ag, da = fa - cross(ld3(sm.sc + 1), ag);
            st3(sm.sc + 22, com);
            st3(sm.sc + 25, (1.0 / mt) * hl);
            st3(sm.sc + 28, (1.0 / mt) * dl);
            st3(sm.sc + 31, hl);
            st3(sm.sc + 34, ha - cross(com, hl));
            st3(sm.sc + 37, dl);
            st3(sm.sc + 40, da - cross(com, dl));
        }
        // ---- frames: placement, local velocity, local classic acceleration
        for (int f = tid; f < z.nf; f += G) {
            const int pj = gP->fparent[f];
            const double* fp = gP->fplace[f];
            double* o = sm.oMf + 12 * f;
            V3 vl = mk(0, 0, 0), w = mk(0, 0, 0), al = mk(0, 0, 0), aa = mk(0, 0, 0);
            if (pj >= 0) {
                const double* pm = sm.oMi + 12 * pj;
                mm3(pm, fp, o);
                st3(o + 9, rmul(pm, ld3(fp + 9)) + ld3(pm + 9));
                vl = ld3(sm.vw + 6 * pj);
                w = ld3(sm.vw + 6 * pj + 3);
                al = ld3(sm.aw + 6 * pj);
                aa = ld3(sm.aw + 6 * pj + 3);
            }
            else {
                for (int k = 0; k < 12; ++k)
                    o[k] = fp[k];
            }
            V3 p = ld3(o + 9);
            V3 wl = rtmul(o, w), vll = rtmul(o, vl - cross(p, w));
            V3 aal = rtmul(o, aa), all = rtmul(o, al - cross(p, aa)) + cross(wl, vll);
            st3(sm.vf + 6 * f, vll);
            st3(sm.vf + 6 * f + 3, wl);
            st3(sm.af + 6 * f, all);
            st3(sm.af + 6 * f + 3, aal);
        }
        Grp<G>::sync();
        (void)flip;
    }

    // local-frame Jacobian column c of frame f (zero if the dof does not move the frame)
    __device__ __forceinline__ void frame_jac_col(const iwbc_problem_t* gP, const Sm& sm, int f, int c, V3& lin, V3& ang)
    {
        const int pj = gP->fparent[f];
        lin = mk(0, 0, 0);
        ang = mk(0, 0, 0);
        if (pj < 0 || !supports(gP, sm.dofj[c], pj))
            return;
        const double* o = sm.oMf + 12 * f;
        V3 sl = ld3(sm.Sw + 6 * c), sa = ld3(sm.Sw + 6 * c + 3);
        ang = rtmul(o, sa);
        lin = rtmul(o, sl - cross(ld3(o + 9), sa));
    }

    // SE3 task right-hand side (a_des - drift), local frame (ex_task.cpp:175-247)
    __device__ inline void se3_rhs(const Sm& sm, int f, const double* kp, const double* kd, const double* ref, double* b6)
    {
        const double* o = sm.oMf + 12 * f;
        // Mref: p = ref[0..2], R column-major ref[3..11]
        double Rr[9];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < 3; ++r)
                Rr[3 * r + c] = ref[3 + 3 * c + r];
        // E = oMf^-1 * Mref
        double Rt[9] = {o[0], o[3], o[6], o[1], o[4], o[7], o[2], o[5], o[8]};
        double Re[9];
        mm3(Rt, Rr, Re);
        V3 pe = rtmul(o, ld3(ref) - ld3(o + 9));
        V3 el, ea;
        log6(Re, pe, el, ea);
        V3 vrl = rtmul(o, ld3(ref + 12)), vra = rtmul(o, ld3(ref + 15));
        V3 arl = rtmul(o, ld3(ref + 18)), ara = rtmul(o, ld3(ref + 21));
        const double* v = sm.vf + 6 * f;
        const double* a = sm.af + 6 * f;
        double e[6] = {el.x, el.y, el.z, ea.x, ea.y, ea.z};
        double vr[6] = {vrl.x, vrl.y, vrl.z, vra.x, vra.y, vra.z};
        double ar[6] = {arl.x, arl.y, arl.z, ara.x, ara.y, ara.z};
#pragma unroll
        for (int i = 0; i < 6; ++i)
            b6[i] = kp[i] * e[i] + kd[i] * (vr[i] - v[i]) + ar[i] - a[i];
    }

    // TaskJointPosVelAccBounds (upstream constructor stores 2*dt)
    __device__ inline void acc_bounds(const iwbc_problem_t* gP, int i, double qi, double dq, double& lb, double& ub)
    {
        const double dt = 2.0 * gP->dt;
        const double qmin = gP->q_lo[i], qmax = gP->q_hi[i], dqmax = gP->v_max[i], ddqmax = gP->v_max[i] / gP->dt;
        const double m = -dq / dt;
        double lbP = -1e10, ubP = 1e10;
        const double q3max = 2 * (qmax - qi - dt * dq) / (dt * dt), q3min = 2 * (qmin - qi - dt * dq) / (dt * dt);
        if (dq <= 0.0) {
            ubP = q3max;
            if (q3min < m)
                lbP = q3min;
            else if (qi != qmin)
                lbP = fmax((dq * dq) / (2 * (qi - qmin)), m);
            else
                lbP = 1e6;
        }
        else {
            lbP = q3min;
            if (q3max > m)
                ubP = q3max;
            else if (qi != qmax)
                ubP = fmin(-(dq * dq) / (2 * (qmax - qi)), m);
            else
                ubP = -1e6;
        }
        const double lbV = (-dqmax - dq) / dt, ubV = (dqmax - dq) / dt;
        const double a2 = 2 * dt * dt;
        const double b1 = 2 * dt * dq + ddqmax * dt * dt, b2 = 2 * dt * dq - ddqmax * dt * dt;
        const double c1 = dq * dq - 2 * ddqmax * (qmax - (qi + dt * dq));
        const double c2 = dq * dq - 2 * ddqmax * ((qi + dt * dq) - qmin);
        const double d1 = b1 * b1 - 2 * a2 * c1, d2 = b2 * b2 - 2 * a2 * c2;
        const double ddq1 = d1 >= 0 ? (-b1 + sqrt(d1)) / a2 : m;
        const double ddq2 = d2 >= 0 ? (-b2 - sqrt(d2)) / a2 : m;
        const double ubVia = fmax(ddq1, m), lbVia = fmin(ddq2, m);
        lb = fmax(fmax(lbP, lbVia), fmax(lbV, -ddqmax));
        ub = fmin(fmin(ubP, ubVia), fmin(ubV, ddqmax));
        if (ub < lb) {
            if (ub == ubP)
                lb = ub;
            else
                ub = lb;
        }
    }

    // =================================================================================== k_asm
    // M(r, k).  msym 0: full nv x ldM.  1: strict lower triangle only, row r at M + r ldM (k < r), diagonal in Md -- the fast pass keeps
    // M in the unused upper triangle of its column-major augmented matrix (ldM = that matrix's leading dimension).  2: packed lower
    // triangle, row r at M + r (r + 1) / 2 (k <= r) -- the continuation's own copy (its R overwrites the augmented matrix).
    __device__ __forceinline__ double mass_at(const Sm& sm, const Sizes& z, int r, int k)
    {
        if (z.msym == 0)
            return sm.M[r * z.ldM + k];
        const int hi = r > k ? r : k, lo = r > k ? k : r;
        if (z.msym == 1)
            return hi == lo ? sm.Md[hi] : sm.M[hi * z.ldM + lo];
        return sm.M[hi * (hi + 1) / 2 + lo];
    }

    // Fills: Jft / cb (contacts), lb / ub (acceleration bounds), A / b (dense level-1 rows, scaled
    // by sqrt(w)), then H (lower triangle, in sm.J) and g.
    template <int G>
    __device__ __forceinline__ void k_asm(const iwbc_problem_t* __restrict__ gP, const Sm& sm, const Sizes& z, int tid, int& flip)
    {
        const int nv = z.nv, n = z.n, na = z.P->na, nu = nv - na;
        const Shape& S = *z.S;
        // ---- contact motion rows
        for (int c = 0; c < z.P->ncontacts; ++c) {
            if (S.cidx[c] < 0)
                continue;
            const iwbc_contact_t& C = z.P->contacts[c];
            for (int k = tid; k < nv; k += G) {
                V3 l, a;
                frame_jac_col(gP, sm, C.frame, k, l, a);
                double* J = sm.Jft + (c * 6) * nv + k;
                J[0] = l.x, J[nv] = l.y, J[2 * nv] = l.z, J[3 * nv] = a.x, J[4 * nv] = a.y, J[5 * nv] = a.z;
            }
            if (tid == (G - 1 - c) % G)
                se3_rhs(sm, C.frame, C.kp, C.kd, sm.ref + C.ref_off, sm.cb + 6 * c);
        }
        // ---- acceleration bounds
        for (int i = tid; i < na; i += G)
            acc_bounds(gP, i, sm.qs[z.nq - na + i], sm.vs[nv - na + i], sm.lb[i], sm.ub[i]);
        // ---- g = 0 ; H diagonal part handled later
        for (int k = tid; k < n; k += G)
            sm.g[k] = 0.0;
        // ---- dense level-1 rows, in HQP order
        int row = 0;
        for (int bl = 0; bl < z.P->nl1; ++bl) {
            if (z.P->l1[bl].kind != IWBC_B_TASK)
                continue;
            const iwbc_task_t& T = z.P->tasks[z.P->l1[bl].index];
            const double sw = sqrt(T.weight);
            if (T.type == IWBC_T_SE3) {
                if (tid == (row + 7) % G) {
                    double b6[6];
                    se3_rhs(sm, T.frame, T.kp, T.kd, sm.ref + T.ref_off, b6);
                    int idx = 0;
                    for (int i = 0; i < 6; ++i)
                        if (T.mask[i])
                            sm.b[row + idx++] = sw * b6[i];
                }
                for (int k = tid; k < nv; k += G) {
                    V3 l, a;
                    frame_jac_col(gP, sm, T.frame, k, l, a);
                    double col[6] = {l.x, l.y, l.z, a.x, a.y, a.z};
                    int idx = 0;
                    for (int i = 0; i < 6; ++i)
                        if (T.mask[i])
                            sm.A[(row + idx++) * nv + k] = sw * col[i];
                }
                row += T.nrows;
            }
            else if (T.type == IWBC_T_COM) {
                const double* r = sm.ref + T.ref_off;
                const double imt = 1.0 / sm.sc[0];
                if (tid == 0) {
                    int idx = 0;
                    for (int i = 0; i < 3; ++i)
                        if (T.mask[i]) {
                            double a_des = -T.kp[i] * (sm.sc[22 + i] - r[i]) - T.kd[i] * (sm.sc[25 + i] - r[3 + i]) + r[6 + i];
                            sm.b[row + idx++] = sw * (a_des - sm.sc[28 + i]);
                        }
                }
                for (int k = tid; k < nv; k += G) {
                    int idx = 0;
                    for (int i = 0; i < 3; ++i)
                        if (T.mask[i])
                            sm.A[(row + idx++) * nv + k] = sw * imt * sm.Fc[6 * k + i]; // Jcom = Ag_lin / m
                }
                row += T.nrows;
            }
            else if (T.type == IWBC_T_MOMENTUM) {
                const double* r = sm.ref + T.ref_off;
                const V3 com = ld3(sm.sc + 22);
                if (tid == 0) {
                    int idx = 0;
                    for (int i = 0; i < 6; ++i)
                        if (T.mask[i]) {
                            double dL = -T.kp[i] * (sm.sc[31 + i] - r[6 + i]) + r[12 + i]; // L = Ag v = hg
                            sm.b[row + idx++] = sw * (dL - sm.sc[37 + i]);
                        }
                }
                for (int k = tid; k < nv; k += G) {
                    V3 Fl = ld3(sm.Fc + 6 * k), Fa = ld3(sm.Fc + 6 * k + 3) - cross(com, ld3(sm.Fc + 6 * k));
                    double col[6] = {Fl.x, Fl.y, Fl.z, Fa.x, Fa.y, Fa.z};
                    int idx = 0;
                    for (int i = 0; i < 6; ++i)
                        if (T.mask[i])
                            sm.A[(row + idx++) * nv + k] = sw * col[i];
                }
                row += T.nrows;
            }
            else if (T.type == IWBC_T_SELFCOL) {
                // per avoided frame a: gradient (3) -> sc[44 + 4a .. +2], and the scalar B contribution -> sc[44 + 4a + 3]
                const int f0 = T.frame;
                const int lwa = !z.P->sc_jacobian_world;
                for (int a = tid; a < T.nav; a += G) {
                    const int fa = T.av_frame[a];
                    const double* o0 = sm.oMf + 12 * f0;
                    const double* oa = sm.oMf + 12 * fa;
                    V3 diff = ld3(o0 + 9) - ld3(oa + 9);
                    double sq = dot(diff, diff), nrm = sqrt(sq);
                    double aa = T.av_r0[a] + T.radius;
                    V3 drift = ld3(sm.af + 6 * f0) - ld3(sm.af + 6 * fa);
                    // J v of each frame: world linear velocity of the frame origin (LWA) or of the origin-coincident point (WORLD)
                    auto point_vel = [&](int f, const double* o) {
                        int pj = gP->fparent[f];
                        if (pj < 0)
                            return mk(0, 0, 0);
                        V3 vl = ld3(sm.vw + 6 * pj), w = ld3(sm.vw + 6 * pj + 3);
                        return lwa ? vl + cross(w, ld3(o + 9)) : vl;
                    };
                    V3 Jv = point_vel(f0, o0) - point_vel(fa, oa);
                    const double m = T.sc_m, k = T.sc_k;
                    double x = k * (nrm - aa + T.sc_sp);
                    double e_p = exp(-x);
                    double l1 = log1p(e_p);            // pow(1 + e_p, p) = exp(p * log1p(e_p))
                    double pw0 = exp(-m * l1);         // (1+e_p)^-m
                    double pw1 = exp((-m - 1) * l1);   // (1+e_p)^(-m-1)
                    double pw2 = exp((-m - 2) * l1);   // (1+e_p)^(-m-2)
                    double Cv = 1. - pw0;
                    double gs = -1. / nrm * k * m * e_p * pw1;
                    double h = 1. / sq * k * k * (-m - 1.) * m * exp(-2 * x) * pw2 + 1. / sq * k * k * m * e_p * pw1
                        + 1. / (nrm * sqrt(nrm)) * k * m * e_p * pw1;
                    V3 grad = gs * diff;
                    V3 HJv = (h * dot(diff, Jv)) * diff + gs * Jv;
                    double Bc = -(dot(HJv, Jv) + dot(grad, T.kd[0] * Jv - drift) + T.kp[0] * Cv);
                    st3(sm.sc + 44 + 4 * a, grad);
                    sm.sc[44 + 4 * a + 3] = Bc;
                }
                Grp<G>::sync();
                if (tid == 0) {
                    double Bv = 0;
                    for (int a = 0; a < T.nav; ++a)
                        Bv += sm.sc[44 + 4 * a + 3];
                    sm.b[row] = sw * Bv;
                }
                for (int k = tid; k < nv; k += G) {
                    const int jk = sm.dofj[k];
                    V3 sl = ld3(sm.Sw + 6 * k), sa = ld3(sm.Sw + 6 * k + 3);
                    auto col = [&](int f) {
                        int pj = gP->fparent[f];
                        if (pj < 0 || !supports(gP, jk, pj))
                            return mk(0, 0, 0);
                        return lwa ? sl + cross(sa, ld3(sm.oMf + 12 * f + 9)) : sl;
                    };
                    V3 c0 = col(f0);
                    double acc = 0;
                    for (int a = 0; a < T.nav; ++a)
                        acc += dot(ld3(sm.sc + 44 + 4 * a), c0 - col(T.av_frame[a]));
                    sm.A[row * nv + k] = sw * acc;
                }
                Grp<G>::sync(); // sc is reused by the next self-collision task
                row += 1;
            }
        }
        Grp<G>::sync();
        // ---- H (lower triangle incl. diagonal) into sm.J (column-major, ld = ldJ): H[i][j], i >= j
        const int ld = z.ldJ;
        const int R = row;
        // dv-dv block from the dense rows
        for (int e = tid; e < nv * (nv + 1) / 2; e += G) {
            // unrank (i >= j) from e: row-wise packing of the lower triangle
            int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (i * (i + 1) / 2 > e)
                --i;
            while ((i + 1) * (i + 2) / 2 <= e)
                ++i;
            const int j = e - i * (i + 1) / 2;
            double s0 = 0, s1 = 0;
            int r = 0;
            for (; r + 1 < R; r += 2) {
                s0 += sm.A[r * nv + i] * sm.A[r * nv + j];
                s1 += sm.A[(r + 1) * nv + i] * sm.A[(r + 1) * nv + j];
            }
            if (r < R)
                s0 += sm.A[r * nv + i] * sm.A[r * nv + j];
            sm.J[i + j * ld] = s0 + s1;
        }
        // g (dv part) = -A' b
        for (int k = tid; k < nv; k += G) {
            double s = 0;
            for (int r = 0; r < R; ++r)
                s += sm.A[r * nv + k] * sm.b[r];
            sm.g[k] = -s;
        }
        // force block and cross block: zero, then the constant force-regularisation blocks
        for (int e = tid; e < z.n * z.n - nv * nv; e += G) {
            // enumerate entries (i >= nv, any j <= i)
            int i = nv, rem = e;
            // rows i have (i + 1) entries each in the lower triangle
            while (rem >= i + 1 && i < n) {
                rem -= i + 1;
                ++i;
            }
            if (i < n)
                sm.J[i + rem * ld] = 0.0;
        }
        Grp<G>::sync();
        for (int bl = 0; bl < z.P->nl1; ++bl) {
            const iwbc_block_t& Bk = z.P->l1[bl];
            if (Bk.kind == IWBC_B_CONTACT_REG) {
                const int c = Bk.index;
                if (S.cidx[c] < 0)
                    continue;
                const iwbc_contact_t& C = z.P->contacts[c];
                const int o = nv + S.cidx[c];
                // f_ref and the installed weight vector come with the instance's reference row (iwbc_problem.h: contact block)
                const double* fref = sm.ref + C.ref_off + IWBC_REF_CONTACT_FREF;
                int wv = (int)sm.ref[C.ref_off + IWBC_REF_CONTACT_WREG];
                if (wv < 0 || wv >= z.P->nwreg)
                    wv = 0;
                const double* wreg = z.P->wreg_variants[wv];
                for (int e = tid; e < 78; e += G) { // 12*13/2 lower entries
                    int i = 0, rem = e;
                    while (rem >= i + 1) {
                        rem -= i + 1;
                        ++i;
                    }
                    const int j = rem;
                    double s = 0;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
                        s += (wreg[r] * C.T[r * 12 + i]) * (wreg[r] * C.T[r * 12 + j]);
                    sm.J[(o + i) + (o + j) * ld] += C.w_forcereg * s;
                }
                for (int k = tid; k < 12; k += G) {
                    double s = 0;
#pragma unroll
                    for (int r = 0; r < 6; ++r)
                        s += (wreg[r] * C.T[r * 12 + k]) * (wreg[r] * fref[r]);
                    sm.g[o + k] -= C.w_forcereg * s;
                }
            }
            else if (z.P->tasks[Bk.index].type == IWBC_T_COP) {
                // TaskCopEquality as a level-1 cost (tasks.cpp:156-178): w || sum_c sum_j (p_cj - cop_ref) (n' f_cj) ||^2 over the ACTIVE
                // contacts' four points (p_cj: world position of contact point j of contact c, f_cj: its three force coefficients).
                // Force columns only, zero right-hand side: it makes the force block of H depend on the state, so the fast path
                // (constant force-block factor) never takes a problem with this task.
                const iwbc_task_t& T = z.P->tasks[Bk.index];
                const double* r = sm.ref + T.ref_off; // cop reference (3), contact normal (3)
                // entry (row d, force column kf of the QP) of the task matrix: (p - ref)[d] * normal[kf % 3]
                auto Adk = [&](int d, int kf) -> double {
                    int c = 0;
                    for (int cc = 0; cc < z.P->ncontacts; ++cc)
                        if (S.cidx[cc] >= 0 && kf >= S.cidx[cc] && kf < S.cidx[cc] + 12)
                            c = cc;
                    const int kk = kf - S.cidx[c], j = kk / 3, a = kk - 3 * j;
                    const iwbc_contact_t& C = z.P->contacts[c];
                    // contact point j in the contact frame, read back from the force generator T = [I ; skew(P_j)] (rows 3..5)
                    const V3 Pl = mk(C.T[5 * 12 + 3 * j + 1], C.T[3 * 12 + 3 * j + 2], C.T[4 * 12 + 3 * j + 0]);
                    const double* o = sm.oMf + 12 * C.frame;
                    const V3 pw = rmul(o, Pl) + ld3(o + 9);
                    const double pd = d == 0 ? pw.x - r[0] : (d == 1 ? pw.y - r[1] : pw.z - r[2]);
                    return pd * r[3 + a];
                };
                const int nk = n - nv;
                for (int e = tid; e < nk * (nk + 1) / 2; e += G) {
                    int a = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                    while (a * (a + 1) / 2 > e)
                        --a;
                    while ((a + 1) * (a + 2) / 2 <= e)
                        ++a;
                    const int b = e - a * (a + 1) / 2;
                    double s = 0;
                    for (int d = 0; d < 3; ++d)
                        s += Adk(d, a) * Adk(d, b);
                    sm.J[(nv + a) + (nv + b) * ld] += T.weight * s;
                }
            }
            else if (z.P->tasks[Bk.index].type == IWBC_T_TORQUE) {
                // TaskActuationEquality as a level-1 cost (tasks.cpp:227-271): w || diag(s) S (M_a dv - J_a^T f + h_a - tau_ref) ||^2.
                // Row i over [dv ; f]: D(i, k) = M(nu + i, k) for k < nv, -(T_c^T Jft_c)[k'][nu + i] for the force columns; it couples
                // dv and f, so the fast path (block-diagonal H) never takes a problem with this task.
                const iwbc_task_t& T = z.P->tasks[Bk.index];
                const double* r = sm.ref + T.ref_off; // tau_ref (na), weight vector (na)
                auto Dik = [&](int i, int k) -> double {
                    if (k < nv)
                        return mass_at(sm, z, nu + i, k);
                    int c = 0;
                    for (int cc = 0; cc < z.P->ncontacts; ++cc)
                        if (S.cidx[cc] >= 0 && k - nv >= S.cidx[cc] && k - nv < S.cidx[cc] + 12)
                            c = cc;
                    const int kk = k - nv - S.cidx[c];
                    const double* Tm = z.P->contacts[c].T;
                    double s = 0;
#pragma unroll
                    for (int a = 0; a < 6; ++a)
                        s += Tm[a * 12 + kk] * sm.Jft[(c * 6 + a) * nv + nu + i];
                    return -s;
                };
                for (int e = tid; e < n * (n + 1) / 2; e += G) {
                    int a = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
                    while (a * (a + 1) / 2 > e)
                        --a;
                    while ((a + 1) * (a + 2) / 2 <= e)
                        ++a;
                    const int b = e - a * (a + 1) / 2;
                    double s = 0;
                    for (int i = 0; i < na; ++i)
                        if (T.mask[i])
                            s += (r[na + i] * Dik(i, a)) * (r[na + i] * Dik(i, b));
                    sm.J[a + b * ld] += T.weight * s;
                }
                for (int k = tid; k < n; k += G) {
                    double s = 0;
                    for (int i = 0; i < na; ++i)
                        if (T.mask[i])
                            s += (r[na + i] * Dik(i, k)) * (r[na + i] * (r[i] - sm.nle[nu + i]));
                    sm.g[k] -= T.weight * s;
                }
            }
            else {
                const iwbc_task_t& T = z.P->tasks[Bk.index];
                if (T.type != IWBC_T_POSTURE)
                    continue;
                const double* r = sm.ref + T.ref_off;
                for (int i = tid; i < na; i += G) {
                    if (!T.mask[i])
                        continue;
                    double a_des = -T.kp[0] * (sm.qs[z.nq - na + i] - r[i]) - T.kd[0] * (sm.vs[nv - na + i] - r[na + i]) + r[2 * na + i];
                    sm.J[(nu + i) + (nu + i) * ld] += T.weight;
                    sm.g[nu + i] -= T.weight * a_des;
                }
            }
            Grp<G>::sync();
        }
        for (int k = tid; k < n; k += G)
            sm.J[k + k * ld] += z.P->hess_reg;
        Grp<G>::sync();
        (void)flip;
    }

    // =================================================================================== k_qp
    // Implicit constraint rows of the fused problem.
    struct FusedRows {
        // np <- CE.row(i) ; returns ce0_i
        template <int G>
        static __device__ double load_ce(const Sm& sm, const Sizes& z, int i, int tid)
        {
            const Shape& S = *z.S;
            const int nv = z.nv, n = z.n;
            int b = 0;
            while (b + 1 < S.nce && i >= S.ce[b + 1].off)
                ++b;
            const int r = i - S.ce[b].off;
            if (S.ce[b].kind == 0) { // base dynamics row r: [M_u | -J_u^T], ce0 = h_u
                for (int k = tid; k < n; k += G) {
                    double val;
                    if (k < nv)
                        val = mass_at(sm, z, r, k);
                    else {
                        // -(T^T Jft_c)[kk][r]
                        int c = 0;
                        for (int cc = 0; cc < z.P->ncontacts; ++cc)
                            if (S.cidx[cc] >= 0 && k - nv >= S.cidx[cc] && k - nv < S.cidx[cc] + 12)
                                c = cc;
                        const int kk = k - nv - S.cidx[c];
                        const double* T = z.P->contacts[c].T;
                        double s = 0;
#pragma unroll
                        for (int a = 0; a < 6; ++a)
                            s += T[a * 12 + kk] * sm.Jft[(c * 6 + a) * nv + r];
                        val = -s;
                    }
                    sm.np[k] = val;
                }
                return sm.nle[r];
            }
            const int c = S.ce[b].index;
            for (int k = tid; k < n; k += G)
                sm.np[k] = k < nv ? sm.Jft[(c * 6 + r) * nv + k] : 0.0;
            return -sm.cb[6 * c + r];
        }
        // wrench of each active contact w_c = T f_c -> sm.cb[6*ncontacts + 6c + a]
        template <int G>
        static __device__ void wrenches(const Sm& sm, const Sizes& z, int tid)
        {
            const Shape& S = *z.S;
            for (int e = tid; e < 6 * z.P->ncontacts; e += G) {
                const int c = e / 6, a = e - 6 * c;
                double s = 0;
                if (S.cidx[c] >= 0) {
                    const double* T = z.P->contacts[c].T + a * 12;
                    const double* f = sm.x + z.nv + S.cidx[c];
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        s += T[k] * f[k];
                }
                sm.cb[6 * z.P->ncontacts + e] = s;
            }
        }
        // tau~_i = M_a[i] dv - J_a[:, i]^T f  (needs wrenches())
        static __device__ double act_row_value(const Sm& sm, const Sizes& z, int i)
        {
            const Shape& S = *z.S;
            const int nv = z.nv, nu = nv - z.P->na;
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            if (z.msym == 0) {
                const double* Mr = sm.M + (nu + i) * z.ldM;
                int k = 0;
                for (; k + 3 < nv; k += 4) {
                    s0 += Mr[k] * sm.x[k];
                    s1 += Mr[k + 1] * sm.x[k + 1];
                    s2 += Mr[k + 2] * sm.x[k + 2];
                    s3 += Mr[k + 3] * sm.x[k + 3];
                }
                for (; k < nv; ++k)
                    s0 += Mr[k] * sm.x[k];
            }
            else {
                // one stored half: the row's own part (k < r, contiguous), the diagonal, then column r of the rows below
                const int r = nu + i;
                const double* Mr = z.msym == 1 ? sm.M + r * z.ldM : sm.M + r * (r + 1) / 2;
                int k = 0;
                for (; k + 3 < r; k += 4) {
                    s0 += Mr[k] * sm.x[k];
                    s1 += Mr[k + 1] * sm.x[k + 1];
                    s2 += Mr[k + 2] * sm.x[k + 2];
                    s3 += Mr[k + 3] * sm.x[k + 3];
                }
                for (; k < r; ++k)
                    s0 += Mr[k] * sm.x[k];
                s1 += (z.msym == 1 ? sm.Md[r] : Mr[r]) * sm.x[r];
                if (z.msym == 1) {
                    for (k = r + 1; k + 1 < nv; k += 2) {
                        s2 += sm.M[k * z.ldM + r] * sm.x[k];
                        s3 += sm.M[(k + 1) * z.ldM + r] * sm.x[k + 1];
                    }
                    if (k < nv)
                        s2 += sm.M[k * z.ldM + r] * sm.x[k];
                }
                else {
                    for (k = r + 1; k + 1 < nv; k += 2) {
                        s2 += sm.M[k * (k + 1) / 2 + r] * sm.x[k];
                        s3 += sm.M[(k + 1) * (k + 2) / 2 + r] * sm.x[k + 1];
                    }
                    if (k < nv)
                        s2 += sm.M[k * (k + 1) / 2 + r] * sm.x[k];
                }
            }
            double s = (s0 + s1) + (s2 + s3);
            for (int c = 0; c < z.P->ncontacts; ++c) {
                if (S.cidx[c] < 0)
                    continue;
                const double* w = sm.cb + 6 * z.P->ncontacts + 6 * c;
                const double* Jc = sm.Jft + (c * 6) * nv + nu + i;
                const double c0 = Jc[0] * w[0] + Jc[nv] * w[1], c1 = Jc[2 * nv] * w[2] + Jc[3 * nv] * w[3], c2 = Jc[4 * nv] * w[4] + Jc[5 * nv] * w[5];
                s -= (c0 + c1) + c2;
            }
            return s;
        }
        // s = CI x + ci0 for all rows
        template <int G>
        static __device__ void eval_s(const iwbc_problem_t* gP, const Sm& sm, const Sizes& z, int tid)
        {
            const Shape& S = *z.S;
            const int nv = z.nv, na = z.P->na, nu = nv - na;
            wrenches<G>(sm, z, tid);
            Grp<G>::sync();
            for (int b = 0; b < S.nci; ++b) {
                const CiBlock& B = S.ci[b];
                if (B.kind == 0) {
                    for (int i = tid; i < na; i += G) {
                        const double xi = sm.x[nu + i];
                        sm.s[B.off + i] = xi - sm.lb[i];
                        sm.s[B.off + na + i] = -xi + sm.ub[i];
                    }
                }
                else if (B.kind == 1) {
                    for (int i = tid; i < na; i += G) {
                        const double t = act_row_value(sm, z, i);
                        const double tm = gP->tau_max[i], h = sm.nle[nu + i];
                        sm.s[B.off + i] = t - (-tm - h);
                        sm.s[B.off + na + i] = -t + (tm - h);
                    }
                }
                else {
                    const iwbc_contact_t& C = z.P->contacts[B.index];
                    const double* f = sm.x + nv + S.cidx[B.index];
                    for (int r = tid; r < 17; r += G) {
                        double y = 0;
#pragma unroll
                        for (int k = 0; k < 12; ++k)
                            y += C.B[r * 12 + k] * f[k];
                        sm.s[B.off + r] = y - C.lb[r];
                        sm.s[B.off + 17 + r] = -y + C.ub[r];
                    }
                }
            }
        }
        // ci0(ip) of the implicit rows
        static __device__ double ci0_of(const iwbc_problem_t* gP, const Sm& sm, const Sizes& z, int ip)
        {
            const Shape& S = *z.S;
            const int nu = z.nv - z.P->na;
            int b = 0;
            while (b + 1 < S.nci && ip >= S.ci[b + 1].off)
                ++b;
            const CiBlock& B = S.ci[b];
            int r = ip - B.off;
            const bool neg = r >= B.rows;
            if (neg)
                r -= B.rows;
            if (B.kind == 0)
                return neg ? sm.ub[r] : -sm.lb[r];
            if (B.kind == 1)
                return neg ? (gP->tau_max[r] - sm.nle[nu + r]) : (gP->tau_max[r] + sm.nle[nu + r]);
            return neg ? z.P->contacts[B.index].ub[r] : -z.P->contacts[B.index].lb[r];
        }
        // np <- CI.row(ip)
        template <int G>
        static __device__ void load_ci(const Sm& sm, const Sizes& z, int ip, int tid)
        {
            const Shape& S = *z.S;
            const int nv = z.nv, n = z.n, na = z.P->na, nu = nv - na;
            int b = 0;
            while (b + 1 < S.nci && ip >= S.ci[b + 1].off)
                ++b;
            const CiBlock& B = S.ci[b];
            int r = ip - B.off;
            const double sgn = r < B.rows ? 1.0 : -1.0;
            if (r >= B.rows)
                r -= B.rows;
            for (int k = tid; k < n; k += G) {
                double val = 0;
                if (B.kind == 0)
                    val = (k == nu + r) ? 1.0 : 0.0;
                else if (B.kind == 1) {
                    if (k < nv)
                        val = mass_at(sm, z, nu + r, k);
                    else {
                        int c = 0;
                        for (int cc = 0; cc < z.P->ncontacts; ++cc)
                            if (S.cidx[cc] >= 0 && k - nv >= S.cidx[cc] && k - nv < S.cidx[cc] + 12)
                                c = cc;
                        const int kk = k - nv - S.cidx[c];
                        const double* T = z.P->contacts[c].T;
                        double s = 0;
#pragma unroll
                        for (int a = 0; a < 6; ++a)
                            s += T[a * 12 + kk] * sm.Jft[(c * 6 + a) * nv + nu + r];
                        val = -s;
                    }
                }
                else {
                    const int o = nv + S.cidx[B.index];
                    if (k >= o && k < o + 12)
                        val = z.P->contacts[B.index].B[r * 12 + (k - o)];
                }
                sm.np[k] = sgn * val;
            }
        }
    };

    // dense rows in global memory (the SolverHQPBase-shaped entry point)
    struct DenseRows {
        const double* CE;
        const double* ce0;
        const double* CI;
        const double* ci0;
        int n, me, mi;
    };

    // ---- shared pieces of the active-set method
    // d = J^T np, z = J[:, iq:] d[iq:]
    template <int G>
    __device__ __forceinline__ void compute_d_z(const Sm& sm, int n, int ld, int iq, int tid)
    {
        for (int j = tid; j < n; j += G) {
            const double* col = sm.J + j * ld;
            double s0 = 0, s1 = 0;
            int i = 0;
            for (; i + 1 < n; i += 2) {
                s0 += col[i] * sm.np[i];
                s1 += col[i + 1] * sm.np[i + 1];
            }
            if (i < n)
                s0 += col[i] * sm.np[i];
            sm.d[j] = s0 + s1;
        }
        Grp<G>::sync();
        for (int i = tid; i < n; i += G) {
            double s0 = 0, s1 = 0;
            int j = iq;
            for (; j + 1 < n; j += 2) {
                s0 += sm.J[i + j * ld] * sm.d[j];
                s1 += sm.J[i + (j + 1) * ld] * sm.d[j + 1];
            }
            if (j < n)
                s0 += sm.J[i + j * ld] * sm.d[j];
            sm.z[i] = s0 + s1;
        }
    }

    // add_constraint: Givens sweep that folds d[iq+1..n-1] into d[iq], applied to the columns of J
    // row-parallel.  Returns false when the new constraint is (numerically) dependent.
    template <int G>
    __device__ bool add_constraint(const Sm& sm, int n, int ld, int& iq, double& R_norm, int tid)
    {
        // coefficients: h_j^2 = sum_{k >= j-1} d_k^2 accumulated from the tail (same order as the sequential chain)
        for (int j = iq + 1 + tid; j < n; j += G) {
            double acc = 0;
            for (int k = n - 1; k >= j; --k)
                acc += sm.d[k] * sm.d[k];
            // "ss" operand: d[j] as left by rotation j+1 = sign(d_j) * sqrt(acc) for j < n-1, original d[n-1] otherwise
            const double hj1 = sqrt(acc); // |value of d[j] when rotation j is applied|
            double ssv = (j == n - 1) ? sm.d[j] : (sm.d[j] < 0.0 ? -hj1 : hj1);
            double ccv = sm.d[j - 1];
            const double h = sqrt(acc + ccv * ccv);
            double cc = 1.0, ss = 0.0, xny = 0.0;
            if (h != 0.0) {
                cc = ccv / h;
                ss = ssv / h;
                if (cc < 0.0) {
                    cc = -cc;
                    ss = -ss;
                }
                xny = ss / (1.0 + cc);
            }
            else
                xny = -2.0; // flag: skip (identity)
            sm.gc[j] = cc;
            sm.gs[j] = ss;
            sm.gx[j] = xny;
        }
        Grp<G>::sync();
        // rotate rows
        for (int i = tid; i < n; i += G) {
            double* row = sm.J + i;
            double t2 = row[(n - 1) * ld];
            for (int j = n - 1; j >= iq + 1; --j) {
                const double t1 = row[(j - 1) * ld];
                const double xny = sm.gx[j];
                if (xny == -2.0) { // h == 0: no rotation; column j keeps t2, carry becomes column j-1
                    row[j * ld] = t2;
                    t2 = t1;
                    continue;
                }
                const double cc = sm.gc[j], ss = sm.gs[j];
                const double nj1 = t1 * cc + t2 * ss;
                row[j * ld] = xny * (t1 + nj1) - t2;
                t2 = nj1;
            }
            row[iq * ld] = t2;
        }
        // new d[iq] = signed norm of the tail (what the sequential sweep leaves in d[iq])
        double diq;
        {
            double acc = 0;
            for (int k = n - 1; k >= iq + 1; --k)
                acc += sm.d[k] * sm.d[k];
            const double dq = sm.d[iq];
            if (n - 1 >= iq + 1) {
                const double h = sqrt(acc + dq * dq);
                diq = (acc + dq * dq) == 0.0 ? dq : (dq < 0.0 ? -h : h);
            }
            else
                diq = dq;
        }
        Grp<G>::sync(); // all readers of d are done
        // R column iq <- d[0..iq]
        for (int i = tid; i <= iq; i += G)
            sm.R[Ridx(i, iq)] = (i == iq) ? diq : sm.d[i];
        iq++;
        Grp<G>::sync();
        if (fabs(diq) <= 2.220446049250313e-16 * R_norm)
            return false;
        R_norm = fmax(R_norm, fabs(diq));
        return true;
    }

    // delete_constraint(l): remove active constraint l (position qq >= me); serial on R (small), row-parallel on J
    template <int G>
    __device__ void delete_constraint(const Sm& sm, int n, int ld, int me, int& iq, int l, int tid)
    {
        int qq = me;
        for (int i = me; i < iq; ++i)
            if (sm.act[i] == l) {
                qq = i;
                break;
            }
        Grp<G>::sync();
        if (tid == 0) {
            // shift A, u and the columns of R; the shifted column i (old i+1) has a sub-diagonal entry kept in sub[i]
            for (int i = qq; i < iq - 1; ++i) {
                sm.act[i] = sm.act[i + 1];
                sm.u[i] = sm.u[i + 1];
                for (int k = 0; k <= i; ++k)
                    sm.R[Ridx(k, i)] = sm.R[Ridx(k, i + 1)];
                sm.sub[i] = sm.R[Ridx(i + 1, i + 1)];
            }
            sm.act[iq - 1] = sm.act[iq];
            sm.u[iq - 1] = sm.u[iq];
            sm.act[iq] = 0;
            sm.u[iq] = 0.0;
            const int niq = iq - 1;
            // Givens on rows (j, j+1) of R, coefficients kept for the J update
            for (int j = qq; j < niq; ++j) {
                double cc = sm.R[Ridx(j, j)], ss = sm.sub[j];
                const double a1 = fabs(cc), b1 = fabs(ss);
                double h;
                if (a1 > b1) {
                    double t = b1 / a1;
                    h = a1 * sqrt(1.0 + t * t);
                }
                else if (b1 > a1) {
                    double t = a1 / b1;
                    h = b1 * sqrt(1.0 + t * t);
                }
                else
                    h = a1 * sqrt(2.0);
                if (h == 0.0) {
                    sm.gx[j] = -2.0;
                    continue;
                }
                cc = cc / h;
                ss = ss / h;
                if (cc < 0.0) {
                    sm.R[Ridx(j, j)] = -h;
                    cc = -cc;
                    ss = -ss;
                }
                else
                    sm.R[Ridx(j, j)] = h;
                const double xny = ss / (1.0 + cc);
                for (int k = j + 1; k < niq; ++k) {
                    double t1 = sm.R[Ridx(j, k)];
                    double t2 = sm.R[Ridx(j + 1, k)];
                    double r1 = t1 * cc + t2 * ss;
                    sm.R[Ridx(j, k)] = r1;
                    sm.R[Ridx(j + 1, k)] = xny * (t1 + r1) - t2;
                }
                sm.gc[j] = cc;
                sm.gs[j] = ss;
                sm.gx[j] = xny;
            }
        }
        Grp<G>::sync();
        const int niq = iq - 1;
        for (int i = tid; i < n; i += G) {
            double* row = sm.J + i;
            for (int j = qq; j < niq; ++j) {
                const double xny = sm.gx[j];
                if (xny == -2.0)
                    continue;
                const double t1 = row[j * ld], t2 = row[(j + 1) * ld];
                const double r1 = t1 * sm.gc[j] + t2 * sm.gs[j];
                row[j * ld] = r1;
                row[(j + 1) * ld] = xny * (r1 + t1) - t2;
            }
        }
        iq = niq;
        Grp<G>::sync();
    }

    // r[me..iq-1] = (R^-1 d)[me..iq-1] (bottom rows of the back-substitution; see header comment)
    __device__ __forceinline__ void update_r_tail(const Sm& sm, int me, int iq)
    {
        for (int i = iq - 1; i >= me; --i) {
            double s = sm.d[i];
            for (int j = i + 1; j < iq; ++j)
                s -= sm.R[Ridx(i, j)] * sm.r[j];
            sm.r[i] = s / sm.R[Ridx(i, i)];
        }
    }

    // Cholesky of the lower triangle held in sm.J (in place), then J <- L^-T (upper), c1 = trace(H), c2 = trace(J).
    // Returns false if H is not positive definite.
    template <int G>
    __device__ __forceinline__ bool factor_and_invert(const Sm& sm, int n, int ld, double& c1, double& c2, int tid, int& flip)
    {
        {
            double t[1] = {0};
            for (int k = tid; k < n; k += G)
                t[0] += sm.J[k + k * ld];
            gsum<G, 1>(t, sm.red, flip);
            c1 = t[0];
        }
        // left-looking Cholesky: column k
        bool ok = true;
        for (int k = 0; k < n; ++k) {
            for (int i = k + tid; i < n; i += G) {
                double s0 = sm.J[i + k * ld], s1 = 0;
                int j = 0;
                for (; j + 1 < k; j += 2) {
                    s0 -= sm.J[i + j * ld] * sm.J[k + j * ld];
                    s1 -= sm.J[i + (j + 1) * ld] * sm.J[k + (j + 1) * ld];
                }
                if (j < k)
                    s0 -= sm.J[i + j * ld] * sm.J[k + j * ld];
                sm.d[i] = s0 + s1; // staging: column k is read by everybody as L[k][j], only row k..n-1 of column k is written
            }
            Grp<G>::sync();
            const double dkk = sm.d[k];
            if (!(dkk > 0.0))
                ok = false;
            const double lkk = sqrt(dkk > 0.0 ? dkk : 1.0);
            for (int i = k + tid; i < n; i += G)
                sm.J[i + k * ld] = (i == k) ? lkk : sm.d[i] / lkk;
            if (tid == 0)
                sm.Ldiag[k] = lkk;
            Grp<G>::sync();
        }
        // U = L^-T: thread j builds row j of U (= column j of L^-1) in the strict upper triangle
        for (int j = tid; j < n; j += G) {
            const double ujj = 1.0 / sm.Ldiag[j];
            for (int i = j + 1; i < n; ++i) {
                // x_i = -(sum_{k=j}^{i-1} L[i][k] x_k) / L[i][i],  x_j = ujj, x_k = U[j][k]
                double s0 = sm.J[i + j * ld] * ujj, s1 = 0;
                int k = j + 1;
                for (; k + 1 < i; k += 2) {
                    s0 += sm.J[i + k * ld] * sm.J[j + k * ld];
                    s1 += sm.J[i + (k + 1) * ld] * sm.J[j + (k + 1) * ld];
                }
                if (k < i)
                    s0 += sm.J[i + k * ld] * sm.J[j + k * ld];
                sm.J[j + i * ld] = -(s0 + s1) / sm.Ldiag[i];
            }
        }
        Grp<G>::sync();
        return ok;
    }

    // x = -H^-1 g using U = L^-T:  H^-1 = U U^T  (called before the strict lower triangle of J is cleared:
    // uses only the strict upper part and Ldiag)
    template <int G>
    __device__ __forceinline__ void unconstrained_min(const Sm& sm, int n, int ld, int tid)
    {
        // y = U^T g  (y_j = sum_{i <= j} U[i][j] g_i)
        for (int j = tid; j < n; j += G) {
            double s = sm.g[j] / sm.Ldiag[j];
            for (int i = 0; i < j; ++i)
                s += sm.J[i + j * ld] * sm.g[i];
            sm.d[j] = s;
        }
        Grp<G>::sync();
        // x = -U y  (x_i = -sum_{j >= i} U[i][j] y_j)
        for (int i = tid; i < n; i += G) {
            double s = sm.d[i] / sm.Ldiag[i];
            for (int j = i + 1; j < n; ++j)
                s += sm.J[i + j * ld] * sm.d[j];
            sm.x[i] = -s;
        }
        Grp<G>::sync();
    }

    // finish J: diagonal = 1 / Ldiag, strict lower = 0; returns trace(J)
    template <int G>
    __device__ __forceinline__ double finish_J(const Sm& sm, int n, int ld, int tid, int& flip)
    {
        for (int e = tid; e < n * n; e += G) {
            const int j = e / n, i = e - j * n;
            if (i > j)
                sm.J[i + j * ld] = 0.0;
            else if (i == j)
                sm.J[i + j * ld] = 1.0 / sm.Ldiag[j];
        }
        double t[1] = {0};
        for (int k = tid; k < n; k += G)
            t[0] += 1.0 / sm.Ldiag[k];
        gsum<G, 1>(t, sm.red, flip);
        Grp<G>::sync();
        return t[0];
    }

    template <int G, class ROWS>
    __device__ __forceinline__ int gi_loop(const iwbc_problem_t* gP, const Sm& sm, const Sizes& z, const ROWS* dense, int me, int mi, int tid, int& flip,
        int& iter_out, int& iq_out, double c1, double c2, double R_norm);

    // The Goldfarb-Idnani loop.  ROWS provides the constraint rows.
    // Returns the eiquadprog status mapped to iwbc_status; iq_out = final active-set size.
    template <int G, class ROWS>
    __device__ __forceinline__ int gi_solve(const iwbc_problem_t* gP, const Sm& sm, const Sizes& z, const ROWS* dense, int me, int mi, int tid, int& flip,
        int& iter_out, int& iq_out, long long* cyc = nullptr)
    {
        const int n = z.n, ld = z.ldJ;
        const double eps = 2.220446049250313e-16;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        double c1, c2;
        if (!factor_and_invert<G>(sm, n, ld, c1, c2, tid, flip)) {
            iter_out = 0;
            iq_out = 0;
            return IWBC_STATUS_INFEASIBLE;
        }
        unconstrained_min<G>(sm, n, ld, tid);
        c2 = finish_J<G>(sm, n, ld, tid, flip);
        if (cyc && tid == 0)
            cyc[3] = clock64();
        double R_norm = 1.0;
        int iq = 0, iter = 0;

        // ---- equality constraints
        for (int i = 0; i < me; ++i) {
            double ce0;
            if constexpr (std::is_same<ROWS, FusedRows>::value)
                ce0 = FusedRows::load_ce<G>(sm, z, i, tid);
            else {
                for (int k = tid; k < n; k += G)
                    sm.np[k] = dense->CE[(size_t)i * n + k];
                ce0 = dense->ce0[i];
            }
            Grp<G>::sync();
            compute_d_z<G>(sm, n, ld, iq, tid);
            Grp<G>::sync();
            double t[3] = {0, 0, 0};
            for (int k = tid; k < n; k += G) {
                t[0] += sm.z[k] * sm.z[k];
                t[1] += sm.z[k] * sm.np[k];
                t[2] += sm.np[k] * sm.x[k];
            }
            gsum<G, 3>(t, sm.red, flip);
            double t2 = 0.0;
            if (fabs(t[0]) > eps)
                t2 = (-t[2] - ce0) / t[1];
            for (int k = tid; k < n; k += G)
                sm.x[k] += t2 * sm.z[k];
            if (tid == 0) {
                sm.u[iq] = t2;
                sm.act[i] = -i - 1;
            }
            if (!add_constraint<G>(sm, n, ld, iq, R_norm, tid)) {
                iter_out = 0;
                iq_out = iq;
                return IWBC_STATUS_ERROR; // REDUNDANT_EQUALITIES
            }
        }
        if (cyc && tid == 0)
            cyc[4] = clock64();
        return gi_loop<G, ROWS>(gP, sm, z, dense, me, mi, tid, flip, iter_out, iq_out, c1, c2, R_norm);
    }

    // The inequality iterations of Goldfarb-Idnani, started from the minimiser on the equality constraints:
    // J (n x n, J'HJ = I, first me columns span H^-1 N_eq), R (me x me), x, c1 = trace(H), c2 = trace(L^-T), R_norm given.
    template <int G, class ROWS>
    __device__ __forceinline__ int gi_loop(const iwbc_problem_t* gP, const Sm& sm, const Sizes& z, const ROWS* dense, int me, int mi, int tid, int& flip,
        int& iter_out, int& iq_out, double c1, double c2, double R_norm)
    {
        const int n = z.n, ld = z.ldJ;
        const double eps = 2.220446049250313e-16;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        int iq = me, iter = 0;
        // iai = K \ A ; all constraints eligible
        for (int k = tid; k < mi; k += G)
            sm.iai[k] = k;
        Grp<G>::sync();

        int status = IWBC_STATUS_UNKNOWN;
        int ip = 0;
        bool need_l1 = true; // start at l1
        bool need_l2 = false;
        double sip = 0.0; // s(ip), kept in a register by every thread
        while (true) {
            if (need_l1) {
                need_l1 = false;
                iter++;
                if (iter >= 1000) {
                    status = IWBC_STATUS_MAX_ITER;
                    break;
                }
                for (int k = me + tid; k < iq; k += G)
                    sm.iai[sm.act[k]] = -1;
                // s = CI x + ci0
                if constexpr (std::is_same<ROWS, FusedRows>::value)
                    FusedRows::eval_s<G>(gP, sm, z, tid);
                else {
                    for (int r = tid; r < mi; r += G) {
                        const double* row = dense->CI + (size_t)r * n;
                        double s0 = dense->ci0[r];
                        for (int k = 0; k < n; ++k)
                            s0 += row[k] * sm.x[k];
                        sm.s[r] = s0;
                    }
                }
                Grp<G>::sync();
                double t[1] = {0};
                for (int k = tid; k < mi; k += G) {
                    t[0] += fmin(0.0, sm.s[k]);
                    // iaexcl <- 1 (encoded in iai: excluded rows are marked -2 until the next l1)
                    if (sm.iai[k] == -2)
                        sm.iai[k] = k;
                }
                gsum<G, 1>(t, sm.red, flip);
                if (fabs(t[0]) <= mi * eps * c1 * c2 * 100.0) {
                    status = IWBC_STATUS_OPTIMAL;
                    break;
                }
                // save old values
                for (int k = tid; k < iq; k += G) {
                    sm.u_old[k] = sm.u[k];
                    sm.act_old[k] = sm.act[k];
                }
                for (int k = tid; k < n; k += G)
                    sm.x_old[k] = sm.x[k];
                Grp<G>::sync();
                need_l2 = true;
            }
            if (need_l2) {
                need_l2 = false;
                // most violated eligible constraint; strict '<' starting from 0, first index wins
                double best = 0.0;
                int bi = 0x7fffffff;
                for (int k = tid; k < mi; k += G)
                    if (sm.iai[k] >= 0 && sm.s[k] < best) {
                        best = sm.s[k];
                        bi = k;
                    }
                gargmin<G>(best, bi, sm.red, flip);
                if (best >= 0.0) {
                    status = IWBC_STATUS_OPTIMAL;
                    break;
                }
                ip = bi;
                sip = best;
                if constexpr (std::is_same<ROWS, FusedRows>::value)
                    FusedRows::load_ci<G>(sm, z, ip, tid);
                else
                    for (int k = tid; k < n; k += G)
                        sm.np[k] = dense->CI[(size_t)ip * n + k];
                if (tid == 0) {
                    sm.u[iq] = 0.0;
                    sm.act[iq] = ip;
                }
                Grp<G>::sync();
            }
            // ---- l2a: step direction
            compute_d_z<G>(sm, n, ld, iq, tid);
            Grp<G>::sync();
            if (tid == 0)
                update_r_tail(sm, me, iq);
            double t[2] = {0, 0};
            for (int k = tid; k < n; k += G) {
                t[0] += sm.z[k] * sm.z[k];
                t[1] += sm.z[k] * sm.np[k];
            }
            gsum<G, 2>(t, sm.red, flip); // (its barrier also publishes r)
            if (G <= 32)
                Grp<G>::sync();
            // partial step length t1 and the constraint l to drop
            double t1 = inf;
            int l = 0;
            for (int k = me; k < iq; ++k) {
                const double rk = sm.r[k];
                if (rk > 0.0) {
                    const double tmp = sm.u[k] / rk;
                    if (tmp < t1) {
                        t1 = tmp;
                        l = sm.act[k];
                    }
                }
            }
            const double t2 = fabs(t[0]) > eps ? -sip / t[1] : inf;
            const double tt = fmin(t1, t2);
            if (tt >= inf) {
                status = IWBC_STATUS_INFEASIBLE; // eiquadprog UNBOUNDED (dual) = primal infeasible
                break;
            }
            Grp<G>::sync(); // everybody has read u / r / act
            if (t2 >= inf) {
                // step in dual space only
                for (int k = me + tid; k < iq; k += G)
                    sm.u[k] -= tt * sm.r[k];
                if (tid == 0) {
                    sm.u[iq] += tt;
                    sm.iai[l] = l;
                }
                Grp<G>::sync();
                delete_constraint<G>(sm, n, ld, me, iq, l, tid);
                continue; // l2a
            }
            // step in primal and dual space
            for (int k = tid; k < n; k += G)
                sm.x[k] += tt * sm.z[k];
            for (int k = me + tid; k < iq; k += G)
                sm.u[k] -= tt * sm.r[k];
            if (tid == 0)
                sm.u[iq] += tt;
            Grp<G>::sync();
            if (tt == t2) {
                // full step: add ip to the active set
                if (!add_constraint<G>(sm, n, ld, iq, R_norm, tid)) {
                    // degenerate: exclude ip, restore and pick another constraint
                    delete_constraint<G>(sm, n, ld, me, iq, ip, tid);
                    for (int k = tid; k < mi; k += G)
                        if (sm.iai[k] != -2)
                            sm.iai[k] = k;
                    Grp<G>::sync();
                    if (tid == 0)
                        sm.iai[ip] = -2; // iaexcl(ip) = 0
                    for (int k = tid; k < iq; k += G) {
                        sm.act[k] = sm.act_old[k];
                        sm.u[k] = sm.u_old[k];
                    }
                    Grp<G>::sync();
                    for (int k = me + tid; k < iq; k += G)
                        sm.iai[sm.act[k]] = -1;
                    for (int k = tid; k < n; k += G)
                        sm.x[k] = sm.x_old[k];
                    Grp<G>::sync();
                    need_l2 = true;
                    continue;
                }
                if (tid == 0)
                    sm.iai[ip] = -1;
                Grp<G>::sync();
                need_l1 = true;
                continue;
            }
            // partial step: drop constraint l, update s(ip), back to l2a
            if (tid == 0)
                sm.iai[l] = l;
            Grp<G>::sync();
            delete_constraint<G>(sm, n, ld, me, iq, l, tid);
            {
                double tv[1] = {0};
                for (int k = tid; k < n; k += G)
                    tv[0] += sm.np[k] * sm.x[k];
                gsum<G, 1>(tv, sm.red, flip);
                double ci0;
                if constexpr (std::is_same<ROWS, FusedRows>::value)
                    ci0 = FusedRows::ci0_of(gP, sm, z, ip);
                else
                    ci0 = dense->ci0[ip];
                sip = tv[0] + ci0;
                if (tid == 0)
                    sm.s[ip] = sip;
                Grp<G>::sync();
            }
        }
        iter_out = iter;
        iq_out = iq;
        return status;
    }

    // delete_constraint for the CTA-wide loop: the same update (drop column qq of R, restore the triangle by Givens rotations of
    // rows (j, j+1), rotate the matching columns of J) with the serial part spread out -- rows of R shift in parallel (one thread
    // per row), the rotations of R run in warp 0 with one lane per column (coefficients broadcast by shuffles; one rsqrt and
    // a Newton reciprocal per rotation), act / u shift through registers.  The caller has synchronised.
    template <int G>
    __device__ void delete_constraint_cta(const Sm& sm, int n, int ld, int me, int& iq, int l, int tid)
    {
        int qq = me;
        for (int i = me; i < iq; ++i)
            if (sm.act[i] == l) {
                qq = i;
                break;
            }
        const int niq = iq - 1;
        if (niq - qq > 32) { // (more columns than lanes: the generic form)
            __syncthreads();
            delete_constraint<G>(sm, n, ld, me, iq, l, tid);
            return;
        }
        // ---- A: shift.  Row k of R: R(k, i) <- R(k, i+1); the old diagonal of column k becomes the sub-diagonal of column k-1
        int a_next = 0;
        double u_next = 0.0;
        if (tid >= qq && tid < iq) {
            a_next = sm.act[tid + 1];
            u_next = sm.u[tid + 1];
        }
        for (int k = tid; k < iq; k += G) {
            if (k >= qq + 1)
                sm.sub[k - 1] = sm.R[Ridx(k, k)];
            for (int i = (qq > k ? qq : k); i < niq; ++i)
                sm.R[Ridx(k, i)] = sm.R[Ridx(k, i + 1)];
        }
        __syncthreads();
        if (tid >= qq && tid < iq) {
            sm.act[tid] = a_next;
            sm.u[tid] = u_next;
        }
        if (tid == iq) {
            sm.act[iq] = 0;
            sm.u[iq] = 0.0;
        }
        // ---- B: warp 0, lane c <-> column qq + c
        if (tid < 32) {
            const int k = qq + tid;
            for (int j = qq; j < niq; ++j) {
                double cc = 1.0, ss = 0.0, xny = -2.0;
                if (k == j) {
                    const double a = sm.R[Ridx(j, j)], b2 = sm.sub[j];
                    const double hh = a * a + b2 * b2;
                    if (hh != 0.0) {
                        const double ih = rsqrt(hh), h = hh * ih;
                        cc = a * ih;
                        ss = b2 * ih;
                        if (cc < 0.0) {
                            sm.R[Ridx(j, j)] = -h;
                            cc = -cc;
                            ss = -ss;
                        }
                        else
                            sm.R[Ridx(j, j)] = h;
                        const double a1 = 1.0 + cc;
                        double rcp = (double)(1.0f / (float)a1);
                        rcp = rcp * (2.0 - a1 * rcp);
                        rcp = rcp * (2.0 - a1 * rcp);
                        xny = ss * rcp;
                    }
                    sm.gc[j] = cc;
                    sm.gs[j] = ss;
                    sm.gx[j] = xny;
                }
                cc = __shfl_sync(0xffffffffu, cc, j - qq);
                ss = __shfl_sync(0xffffffffu, ss, j - qq);
                xny = __shfl_sync(0xffffffffu, xny, j - qq);
                if (k > j && k < niq && xny != -2.0) {
                    const double t1 = sm.R[Ridx(j, k)], t2 = sm.R[Ridx(j + 1, k)];
                    const double r1 = t1 * cc + t2 * ss;
                    sm.R[Ridx(j, k)] = r1;
                    sm.R[Ridx(j + 1, k)] = xny * (t1 + r1) - t2;
                }
            }
        }
        __syncthreads();
        // ---- C: the same rotations on the columns of J, one row per thread
        for (int i = tid; i < n; i += G) {
            double* row = sm.J + i;
            for (int j = qq; j < niq; ++j) {
                const double xny = sm.gx[j];
                if (xny == -2.0)
                    continue;
                const double t1 = row[j * ld], t2 = row[(j + 1) * ld];
                const double r1 = t1 * sm.gc[j] + t2 * sm.gs[j];
                row[j * ld] = r1;
                row[(j + 1) * ld] = xny * (r1 + t1) - t2;
            }
        }
        iq = niq;
        __syncthreads();
    }

    // ---- CTA-wide, latency-trimmed form of gi_loop for the fused problem (G >= 64: the CTA is the group).
    // Same iteration, same tie-breaking, same termination test as gi_loop (eiquadprog); what
    // changes is the schedule -- the loop is a chain of short dependent phases, so its cost is barriers and serial
    // sections, not flops:
    //   * s = CI x + ci0, psi = sum min(0, s) and the most violated eligible row come out of ONE pass and ONE combined
    //     reduction (one two-sided row pair per thread);
    //   * r = R^-1 d (the multipliers' direction) and the ratio test run in the last warp by shuffles while the other
    //     warps compute z = J2 d2; the barrier of the |z|^2 / z'n+ reduction publishes both;
    //   * the contact wrenches of the next pass, the new column of R and the iai update ride along the row rotations of
    //     add_constraint; the suffix norm of d is taken from the rotation coefficients instead of being re-summed by every
    //     thread;  iai is kept incrementally (no per-pass rebuild).
    // Six barriers per added constraint instead of ~15.  Thread-indexed table reads (T, B, tau_max) go through global
    // memory / L1 (gP) instead of the constant bank, which would replay once per distinct address.
    template <int G>
    __device__ int gi_loop_cta(const iwbc_problem_t* __restrict__ gP, const Sm& sm, const Sizes& z, int me, int mi, int tid, int& flip, int& iter_out,
        int& iq_out, double c1, double c2, double R_norm, unsigned long long* prof = nullptr)
    {
        // development aid: SM cycles of the loop's phases as seen by thread 0 (barrier waits included), prof[0..7] =
        // pass (s, psi, pick) | n+ load | d | z, r, ratio test, coefficients | step + rotations | drops | #directions | #adds
        long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long tc = prof ? clock64() : 0;
#define IWBC_LOOP_TICK(slot)                 \
    if (prof && tid == 0) {                  \
        const long long now_ = clock64();    \
        pc[slot] += now_ - tc;               \
        tc = now_;                           \
    }
        static_assert(G >= 64 && G % 32 == 0, "gi_loop_cta: the CTA is the group");
        const Shape& S = *z.S;
        const int n = z.n, ld = z.ldJ, nv = z.nv, na = z.P->na, nu = nv - na, ncon = z.P->ncontacts;
        const double eps = 2.220446049250313e-16;
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        const int lane = tid & 31, lastw = G / 32 - 1, warp = tid >> 5;
        int iq = me, iter = 0;
        int nitems = 0;
        for (int b = 0; b < S.nci; ++b)
            nitems += S.ci[b].rows;
        // wrench of each active contact w_c = T f_c -> sm.cb[6 ncontacts + 6 c + a], by the threads [t0, t0 + 6 ncontacts)
        auto wrenches_by = [&](int t0) {
            const int e = tid - t0;
            if (e >= 0 && e < 6 * ncon) {
                const int c = e / 6, a = e - 6 * c;
                double w = 0;
                if (S.cidx[c] >= 0) {
                    const double* T = gP->contacts[c].T + a * 12;
                    const double* f = sm.x + nv + S.cidx[c];
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        w += T[k] * f[k];
                }
                sm.cb[6 * ncon + e] = w;
            }
        };
        // the wrenches follow a primal step x += t z linearly: w += t (T z_f); no barrier between the step and the next pass
        auto wrenches_step = [&](int t0, double tstep) {
            const int e = tid - t0;
            if (e >= 0 && e < 6 * ncon) {
                const int c = e / 6, a = e - 6 * c;
                if (S.cidx[c] >= 0) {
                    const double* T = gP->contacts[c].T + a * 12;
                    const double* zf = sm.z + nv + S.cidx[c];
                    double w = 0;
#pragma unroll
                    for (int k = 0; k < 12; ++k)
                        w += T[k] * zf[k];
                    sm.cb[6 * ncon + e] += tstep * w;
                }
            }
        };
        const int wt0 = n + 6 * ncon <= G ? n : 0; // wrench threads: past the rows of J when there is room
        for (int k = tid; k < mi; k += G)
            sm.iai[k] = k;
        wrenches_by(wt0);
        __syncthreads();

        int status = IWBC_STATUS_UNKNOWN;
        int ip = 0;
        double sip = 0.0;
        bool at_l1 = true, need_pick = false;
        while (true) {
            if (at_l1 || need_pick) {
                // ---- l1 (+ l2): s, psi, most violated eligible row.  need_pick alone (after a degenerate add): s is unchanged,
                // only the pick is redone with the excluded rows (iai == -2) left out.
                double psi = 0.0, best = 0.0;
                int bi = 0x7fffffff;
                if (at_l1) {
                    iter++;
                    if (iter >= 1000) {
                        status = IWBC_STATUS_MAX_ITER;
                        break;
                    }
                    for (int it = tid; it < nitems; it += G) {
                        int b = 0, r = it;
                        while (r >= S.ci[b].rows) {
                            r -= S.ci[b].rows;
                            ++b;
                        }
                        const CiBlock& B = S.ci[b];
                        double val, lo, hi;
                        if (B.kind == 0) {
                            val = sm.x[nu + r], lo = sm.lb[r], hi = sm.ub[r];
                        }
                        else if (B.kind == 1) {
                            val = FusedRows::act_row_value(sm, z, r);
                            const double tm = gP->tau_max[r], h = sm.nle[nu + r];
                            lo = -tm - h, hi = tm - h;
                        }
                        else {
                            const iwbc_contact_t& C = gP->contacts[B.index];
                            const double* f = sm.x + nv + S.cidx[B.index];
                            val = 0;
#pragma unroll
                            for (int k = 0; k < 12; ++k)
                                val += C.B[r * 12 + k] * f[k];
                            lo = C.lb[r], hi = C.ub[r];
                        }
                        const int k1 = B.off + r, k2 = k1 + B.rows;
                        const double s1 = val - lo, s2 = -val + hi;
                        sm.s[k1] = s1;
                        sm.s[k2] = s2;
                        psi += fmin(0.0, s1) + fmin(0.0, s2);
                        int e1 = sm.iai[k1], e2 = sm.iai[k2];
                        if (e1 == -2) // iaexcl <- 1 at every l1
                            sm.iai[k1] = e1 = k1;
                        if (e2 == -2)
                            sm.iai[k2] = e2 = k2;
                        if (e1 >= 0 && (s1 < best || (s1 == best && s1 < 0.0 && k1 < bi)))
                            best = s1, bi = k1;
                        if (e2 >= 0 && (s2 < best || (s2 == best && s2 < 0.0 && k2 < bi)))
                            best = s2, bi = k2;
                    }
                }
                else {
                    for (int k = tid; k < mi; k += G)
                        if (sm.iai[k] >= 0 && sm.s[k] < best)
                            best = sm.s[k], bi = k;
                }
                // combined reduction: psi (sum) and (best, bi) (arg-min, first index wins)
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    psi += __shfl_xor_sync(0xffffffffu, psi, off);
                    const double ov = __shfl_xor_sync(0xffffffffu, best, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ov < best || (ov == best && oi < bi))
                        best = ov, bi = oi;
                }
                {
                    double* r = sm.red + flip * 16;
                    flip ^= 1;
                    if (lane == 0) {
                        r[3 * warp] = psi;
                        r[3 * warp + 1] = best;
                        r[3 * warp + 2] = (double)bi;
                    }
                    __syncthreads();
                    psi = r[0], best = r[1], bi = (int)r[2];
                    for (int w = 1; w < G / 32; ++w) {
                        psi += r[3 * w];
                        const double ov = r[3 * w + 1];
                        const int oi = (int)r[3 * w + 2];
                        if (ov < best || (ov == best && oi < bi))
                            best = ov, bi = oi;
                    }
                }
                IWBC_LOOP_TICK(0)
                if (at_l1 && fabs(psi) <= mi * eps * c1 * c2 * 100.0) {
                    status = IWBC_STATUS_OPTIMAL;
                    break;
                }
                if (best >= 0.0) {
                    status = IWBC_STATUS_OPTIMAL;
                    break;
                }
                if (at_l1) { // save old values (read only by the degenerate-add path, several barriers later)
                    for (int k = tid; k < iq; k += G) {
                        sm.u_old[k] = sm.u[k];
                        sm.act_old[k] = sm.act[k];
                    }
                    for (int k = tid; k < n; k += G)
                        sm.x_old[k] = sm.x[k];
                }
                at_l1 = false;
                need_pick = false;
                ip = bi;
                sip = best;
                // np <- CI.row(ip)
                {
                    int b = 0;
                    while (b + 1 < S.nci && ip >= S.ci[b + 1].off)
                        ++b;
                    const CiBlock& B = S.ci[b];
                    int r = ip - B.off;
                    const double sgn = r < B.rows ? 1.0 : -1.0;
                    if (r >= B.rows)
                        r -= B.rows;
                    for (int k = tid; k < n; k += G) {
                        double val = 0;
                        if (B.kind == 0)
                            val = (k == nu + r) ? 1.0 : 0.0;
                        else if (B.kind == 1) {
                            if (k < nv)
                                val = mass_at(sm, z, nu + r, k);
                            else {
                                int c = 0;
                                for (int cc = 0; cc < ncon; ++cc)
                                    if (S.cidx[cc] >= 0 && k - nv >= S.cidx[cc] && k - nv < S.cidx[cc] + 12)
                                        c = cc;
                                const int kk = k - nv - S.cidx[c];
                                const double* T = gP->contacts[c].T;
                                double a0 = 0;
#pragma unroll
                                for (int a = 0; a < 6; ++a)
                                    a0 += T[a * 12 + kk] * sm.Jft[(c * 6 + a) * nv + nu + r];
                                val = -a0;
                            }
                        }
                        else {
                            const int o = nv + S.cidx[B.index];
                            if (k >= o && k < o + 12)
                                val = gP->contacts[B.index].B[r * 12 + (k - o)];
                        }
                        sm.np[k] = sgn * val;
                    }
                }
                if (tid == 0) {
                    sm.u[iq] = 0.0;
                    sm.act[iq] = ip;
                }
                __syncthreads();
                IWBC_LOOP_TICK(1)
            }
            // ---- l2a: d = J' n+
            for (int j = tid; j < n; j += G) {
                const double* col = sm.J + j * ld;
                double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                int i = 0;
                for (; i + 3 < n; i += 4) {
                    s0 += col[i] * sm.np[i];
                    s1 += col[i + 1] * sm.np[i + 1];
                    s2 += col[i + 2] * sm.np[i + 2];
                    s3 += col[i + 3] * sm.np[i + 3];
                }
                for (; i < n; ++i)
                    s0 += col[i] * sm.np[i];
                sm.d[j] = (s0 + s1) + (s2 + s3);
            }
            __syncthreads();
            IWBC_LOOP_TICK(2)
            // ---- last warp: r = R^-1 d on the inequality part, ratio test;  everybody: z = J2 d2
            const int m = iq - me;
            const long long ph3 = prof ? clock64() : 0;
            if (warp == lastw) {
                double t1 = inf;
                int l = 0;
                if (m > 0 && m <= 32) {
                    const int i = me + lane;
                    const double invd = lane < m ? 1.0 / sm.R[Ridx(i, i)] : 0.0;
                    double acc = lane < m ? sm.d[i] * invd : 0.0; // row i pre-scaled by 1 / R_ii: the chain per column is shuffle + FMA
                    double ri = 0.0;
                    // four columns per trip, their (scaled) entries loaded before the chain: per column the chain is shuffle + FMA
                    auto col = [&](int jj, double rv) {
                        const double rj = __shfl_sync(0xffffffffu, acc, jj);
                        if (lane == jj)
                            ri = rj;
                        else if (lane < jj)
                            acc = fma(-rv, rj, acc);
                    };
                    int jj = m - 1;
                    for (; jj >= 3; jj -= 4) {
                        const double r0 = lane < jj ? sm.R[Ridx(i, me + jj)] * invd : 0.0;
                        const double r1 = lane < jj - 1 ? sm.R[Ridx(i, me + jj - 1)] * invd : 0.0;
                        const double r2 = lane < jj - 2 ? sm.R[Ridx(i, me + jj - 2)] * invd : 0.0;
                        const double r3 = lane < jj - 3 ? sm.R[Ridx(i, me + jj - 3)] * invd : 0.0;
                        col(jj, r0);
                        col(jj - 1, r1);
                        col(jj - 2, r2);
                        col(jj - 3, r3);
                    }
                    for (; jj >= 0; --jj)
                        col(jj, lane < jj ? sm.R[Ridx(i, me + jj)] * invd : 0.0);
                    double ratio = inf;
                    int li = 0x7fffffff;
                    if (lane < m) {
                        sm.r[i] = ri;
                        if (ri > 0.0) {
                            ratio = sm.u[i] / ri;
                            li = lane;
                        }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const double ov = __shfl_xor_sync(0xffffffffu, ratio, off);
                        const int oi = __shfl_xor_sync(0xffffffffu, li, off);
                        if (ov < ratio || (ov == ratio && oi < li))
                            ratio = ov, li = oi;
                    }
                    if (ratio < inf) {
                        t1 = ratio;
                        l = sm.act[me + li];
                    }
                }
                else if (m > 32) {
                    if (lane == 0) {
                        update_r_tail(sm, me, iq);
                        for (int k = me; k < iq; ++k) {
                            const double rk = sm.r[k];
                            if (rk > 0.0) {
                                const double tmp = sm.u[k] / rk;
                                if (tmp < t1) {
                                    t1 = tmp;
                                    l = sm.act[k];
                                }
                            }
                        }
                    }
                }
                if (lane == 0) {
                    sm.gs[n] = t1;
                    sm.gx[n] = (double)l;
                    if (prof)
                        atomicAdd(prof + 8, (unsigned long long)(clock64() - ph3)); // r = R^-1 d + ratio test (last warp)
                }
            }
            // Givens coefficients that would fold d[iq+1..n-1] into d[iq] if ip is added (speculative: a partial step discards
            // them), by ONE warp -- the second-to-last, idle in the z rows for n <= 64: suffix sums of d_k^2 by a shuffle scan
            // (<= 3 elements per lane), then per element  h_j^2 = S_j + d_{j-1}^2  and the "ss" operand of rotation j, which
            // is what rotation j+1 left in d[j]: sign(d_j) sqrt(S_j).
            if (warp == (lastw > 0 ? lastw - 1 : 0)) {
                const int cnt = n - 1 - iq; // j = iq + 1 + e, e = 0 .. cnt-1
                if (cnt > 0 && cnt <= 96) {
                    const int per = (cnt + 31) >> 5, e0 = lane * per;
                    double q[3] = {0, 0, 0};
#pragma unroll
                    for (int u = 0; u < 3; ++u)
                        if (u < per && e0 + u < cnt) {
                            const double dv = sm.d[iq + 1 + e0 + u];
                            q[u] = dv * dv;
                        }
                    const double tl = (q[0] + q[1]) + q[2];
                    double suf = tl;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const double o = __shfl_down_sync(0xffffffffu, suf, off);
                        if (lane + off < 32)
                            suf += o;
                    }
                    // S_j for this lane's elements (descending), h_j = sqrt(S_j + d_{j-1}^2) by one rsqrt each; sqrt(S_j) is the
                    // h of element j+1 (S_j = S_{j+1} + d_j^2), taken from the neighbouring element / lane instead of a second root
                    double run = suf - tl; // sum over the lanes after this one
                    double hv[3] = {0, 0, 0}, ihv[3] = {0, 0, 0}, hhv[3] = {0, 0, 0};
#pragma unroll
                    for (int u = 2; u >= 0; --u) {
                        if (u < per && e0 + u < cnt) {
                            run += q[u];
                            const double ccv = sm.d[iq + e0 + u];
                            const double hh = run + ccv * ccv;
                            hhv[u] = hh;
                            if (hh != 0.0) {
                                ihv[u] = rsqrt(hh);
                                hv[u] = hh * ihv[u];
                            }
                        }
                    }
                    const double hnext = __shfl_down_sync(0xffffffffu, hv[0], 1); // h of the first element of the next lane
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        if (!(u < per && e0 + u < cnt))
                            continue;
                        const int j = iq + 1 + e0 + u;
                        const double ccv = sm.d[j - 1], dj = sm.d[j];
                        const double sq = (u + 1 < per) ? hv[u < 2 ? u + 1 : 2] : hnext; // sqrt(S_j) = h_{j+1}  (unused for j = n-1)
                        double cc = 1.0, ss = 0.0, xny = -2.0; // -2: h == 0, identity
                        if (hhv[u] != 0.0) {
                            const double ssv = (j == n - 1) ? dj : (dj < 0.0 ? -sq : sq);
                            cc = ccv * ihv[u];
                            ss = ssv * ihv[u];
                            if (cc < 0.0) {
                                cc = -cc;
                                ss = -ss;
                            }
                            // 1 / (1 + cc), 1 <= 1 + cc <= 2: single-precision seed + two Newton steps (full double accuracy) instead of a division
                            const double a = 1.0 + cc;
                            double rcp = (double)(1.0f / (float)a);
                            rcp = rcp * (2.0 - a * rcp);
                            rcp = rcp * (2.0 - a * rcp);
                            xny = ss * rcp;
                        }
                        sm.gc[j] = cc;
                        sm.gs[j] = ss;
                        sm.gx[j] = xny;
                        if (e0 + u == 0) // what the sequential sweep leaves in d[iq]: the signed norm of d[iq..n-1]
                            sm.gc[n] = hhv[u] == 0.0 ? ccv : (ccv < 0.0 ? -hv[u] : hv[u]);
                    }
                }
                else if (cnt > 96) { // (not reached by the robots of etc/: n <= 74)
                    for (int c = lane; c < cnt; c += 32) {
                        const int j = iq + 1 + c;
                        double acc = 0;
                        for (int k = n - 1; k >= j; --k)
                            acc += sm.d[k] * sm.d[k];
                        const double ccv = sm.d[j - 1], dj = sm.d[j];
                        const double hh = acc + ccv * ccv, h = sqrt(hh);
                        double cc = 1.0, ss = 0.0, xny = -2.0;
                        if (hh != 0.0) {
                            const double sq = sqrt(acc);
                            cc = ccv / h;
                            ss = ((j == n - 1) ? dj : (dj < 0.0 ? -sq : sq)) / h;
                            if (cc < 0.0) {
                                cc = -cc;
                                ss = -ss;
                            }
                            xny = ss / (1.0 + cc);
                        }
                        sm.gc[j] = cc;
                        sm.gs[j] = ss;
                        sm.gx[j] = xny;
                        if (c == 0)
                            sm.gc[n] = hh == 0.0 ? ccv : (ccv < 0.0 ? -h : h);
                    }
                }
                else if (lane == 0)
                    sm.gc[n] = sm.d[iq];
                if (prof && lane == 0)
                    atomicAdd(prof + 9, (unsigned long long)(clock64() - ph3)); // rotation coefficients (second-to-last warp)
            }
            double t[2] = {0, 0};
            for (int i = tid; i < n; i += G) {
                double s0 = 0, s1 = 0;
                int j = iq;
                for (; j + 1 < n; j += 2) {
                    s0 += sm.J[i + j * ld] * sm.d[j];
                    s1 += sm.J[i + (j + 1) * ld] * sm.d[j + 1];
                }
                if (j < n)
                    s0 += sm.J[i + j * ld] * sm.d[j];
                const double zi = s0 + s1;
                sm.z[i] = zi;
                t[0] += zi * zi;
                t[1] += zi * sm.np[i];
            }
            if (prof && tid == 0)
                atomicAdd(prof + 10, (unsigned long long)(clock64() - ph3)); // z rows (warp 0)
            gsum<G, 2>(t, sm.red, flip); // (its barrier also publishes r, t1, l)
            IWBC_LOOP_TICK(3)
            pc[6] += 1;
            const double t1 = sm.gs[n];
            const int l = (int)sm.gx[n];
            const double t2 = fabs(t[0]) > eps ? -sip / t[1] : inf;
            const double tt = fmin(t1, t2);
            if (tt >= inf) {
                status = IWBC_STATUS_INFEASIBLE; // eiquadprog UNBOUNDED (dual) = primal infeasible
                break;
            }
            if (t2 >= inf) {
                // step in dual space only, drop l
                for (int k = me + tid; k < iq; k += G)
                    sm.u[k] -= tt * sm.r[k];
                if (tid == 0) {
                    sm.u[iq] += tt;
                    sm.iai[l] = l;
                }
                __syncthreads();
                delete_constraint_cta<G>(sm, n, ld, me, iq, l, tid);
                IWBC_LOOP_TICK(5)
                continue; // l2a
            }
            // step in primal and dual space
            for (int k = tid; k < n; k += G)
                sm.x[k] += tt * sm.z[k];
            wrenches_step(wt0, tt);
            for (int k = me + tid; k < iq; k += G)
                sm.u[k] -= tt * sm.r[k];
            if (tid == 0)
                sm.u[iq] += tt;
            if (tt == t2) {
                // ---- full step: add ip.  The rotation coefficients were computed next to z (published by the reduction's
                // barrier): straight to the row rotations of J, with the new column of R, iai and the next pass's wrenches
                // riding along
                const double diq = sm.gc[n];
                for (int i = tid; i < n; i += G) {
                    double* row = sm.J + i;
                    double t2r = row[(n - 1) * ld];
                    // Four columns per trip, every operand loaded before the first store: the compiler may not move a shared-memory
                    // load above a store it cannot disambiguate, and behind the stores each load would wait for the FMA carry
                    // chain of the previous column.  Branch-free; xny == -2 marks an identity rotation.
                    // (an identity rotation is stored as cc = 1, ss = 0: the carry needs no select, only the stored value does)
                    auto rot = [&](int j, double t1r, double xny, double cc, double ss) {
                        const double nj1 = fma(t2r, ss, t1r * cc); // the carry chain: one FMA per column
                        row[j * ld] = (xny == -2.0) ? t2r : xny * (t1r + nj1) - t2r;
                        t2r = nj1;
                    };
                    int j = n - 1;
                    for (; j - 3 >= iq + 1; j -= 4) {
                        const double ta = row[(j - 1) * ld], tb = row[(j - 2) * ld], tc2 = row[(j - 3) * ld], td = row[(j - 4) * ld];
                        const double xa = sm.gx[j], xb = sm.gx[j - 1], xc = sm.gx[j - 2], xd = sm.gx[j - 3];
                        const double ca = sm.gc[j], cb2 = sm.gc[j - 1], cc2 = sm.gc[j - 2], cd = sm.gc[j - 3];
                        const double sa = sm.gs[j], sb = sm.gs[j - 1], sc2 = sm.gs[j - 2], sd = sm.gs[j - 3];
                        rot(j, ta, xa, ca, sa);
                        rot(j - 1, tb, xb, cb2, sb);
                        rot(j - 2, tc2, xc, cc2, sc2);
                        rot(j - 3, td, xd, cd, sd);
                    }
                    for (; j >= iq + 1; --j)
                        rot(j, row[(j - 1) * ld], sm.gx[j], sm.gc[j], sm.gs[j]);
                    row[iq * ld] = t2r;
                }
                for (int i = G - 1 - tid; i <= iq; i += G) // (from the top thread down: those are idle in the rotations)
                    sm.R[Ridx(i, iq)] = (i == iq) ? diq : sm.d[i];
                if (tid == G - 1)
                    sm.iai[ip] = -1;
                iq++;
                __syncthreads();
                IWBC_LOOP_TICK(4)
                pc[7] += 1;
                if (fabs(diq) <= eps * R_norm) {
                    // degenerate: exclude ip, restore and pick another constraint
                    delete_constraint_cta<G>(sm, n, ld, me, iq, ip, tid);
                    for (int k = tid; k < mi; k += G)
                        if (sm.iai[k] != -2)
                            sm.iai[k] = k;
                    __syncthreads();
                    if (tid == 0)
                        sm.iai[ip] = -2; // iaexcl(ip) = 0
                    for (int k = tid; k < iq; k += G) {
                        sm.act[k] = sm.act_old[k];
                        sm.u[k] = sm.u_old[k];
                    }
                    __syncthreads();
                    for (int k = me + tid; k < iq; k += G)
                        sm.iai[sm.act[k]] = -1;
                    for (int k = tid; k < n; k += G)
                        sm.x[k] = sm.x_old[k];
                    __syncthreads();
                    wrenches_by(wt0);
                    __syncthreads();
                    need_pick = true;
                    continue;
                }
                R_norm = fmax(R_norm, fabs(diq));
                at_l1 = true;
                continue;
            }
            // ---- partial step: drop constraint l, update s(ip), back to l2a
            if (tid == 0)
                sm.iai[l] = l;
            __syncthreads();
            delete_constraint_cta<G>(sm, n, ld, me, iq, l, tid);
            {
                double tv[1] = {0};
                for (int k = tid; k < n; k += G)
                    tv[0] += sm.np[k] * sm.x[k];
                gsum<G, 1>(tv, sm.red, flip);
                sip = tv[0] + FusedRows::ci0_of(gP, sm, z, ip);
                if (tid == 0)
                    sm.s[ip] = sip;
                __syncthreads();
            }
            IWBC_LOOP_TICK(5)
        }
        if (prof && tid == 0)
            for (int k = 0; k < 8; ++k)
                atomicAdd(prof + k, (unsigned long long)pc[k]);
#undef IWBC_LOOP_TICK
        iter_out = iter;
        iq_out = iq;
        return status;
    }

    // =================================================================================== k_dec
    // decode (controller.cpp:250-260) + integrate (:254-256) + command repack (:263-282) + result stores.
    // Scratch: sm.d (v+), sm.x_old (q+).  taut: precomputed M_a dv - J_a' f (or nullptr: evaluated here).
    template <int G>
    __device__ __forceinline__ void k_dec(const iwbc_problem_t* __restrict__ gP, const Sm& sm, const Sizes& z, const IO& io, int inst, int tid, int status, int iters,
        int iq, const double* taut)
    {
        const int nv = z.nv, nq = z.nq, n = z.n, na = z.P->na, nu = nv - na;
        // ---- decode (controller.cpp:250-260) + integrate (:254-256) + command repack (:263-282)
        if (!taut) {
            FusedRows::wrenches<G>(sm, z, tid);
            Grp<G>::sync();
        }
        if (io.dbg_x && inst == 0)
            for (int k = tid; k < n; k += G)
                io.dbg_x[k] = sm.x[k];
        const double dt = z.P->dt;
        if (io.tau)
            for (int i = tid; i < na; i += G)
                io.tau[(size_t)inst * na + i] = sm.nle[nu + i] + (taut ? taut[i] : FusedRows::act_row_value(sm, z, i));
        if (io.dv)
            for (int k = tid; k < nv; k += G)
                io.dv[(size_t)inst * nv + k] = sm.x[k];
        if (io.f)
            for (int k = tid; k < 12 * z.nc; k += G) {
                const int c = k / 12;
                io.f[(size_t)inst * 12 * z.nc + k] = z.S->cidx[c] >= 0 ? sm.x[nv + z.S->cidx[c] + (k - 12 * c)] : 0.0;
            }
        // v+ -> sm.d (nv <= n), q+ -> sm.z region is too small for nq = nv + 1 when nk = 0: use sm.np / sm.gc
        for (int k = tid; k < nv; k += G)
            sm.d[k] = sm.vs[k] + dt * sm.x[k];
        Grp<G>::sync();
        double* qn = sm.x_old; // n + 2 >= nq
        for (int j = tid; j < z.nj; j += G) {
            const int iqx = gP->idx_q[j], iv = gP->idx_v[j];
            if (gP->jtype[j] != IWBC_J_FREEFLYER) {
                qn[iqx] = sm.qs[iqx] + dt * sm.d[iv];
                continue;
            }
            // SE(3) exponential of dt * v+ composed on the right (pinocchio JointModelFreeFlyer::integrate)
            const double* qq = sm.qs + iqx;
            V3 vl = dt * ld3(sm.d + iv), w = dt * ld3(sm.d + iv + 3);
            const double t2 = dot(w, w), t = sqrt(t2);
            double bb, cc, s2, c2;
            if (t < 1.220703125e-4) {
                bb = 0.5 - t2 / 24;
                cc = 1.0 / 6 - t2 / 120;
                s2 = 0.5 - t2 / 48;
                c2 = 1 - t2 / 8;
            }
            else {
                double st, ct;
                sincos(t, &st, &ct);
                bb = (1 - ct) / t2;
                cc = (t - st) / (t2 * t);
                double sh, ch;
                sincos(0.5 * t, &sh, &ch);
                s2 = sh / t;
                c2 = ch;
            }
            V3 wv = cross(w, vl);
            V3 pe = vl + bb * wv + cc * cross(w, wv);
            double R0[9];
            quat_R(qq[3], qq[4], qq[5], qq[6], R0);
            V3 pn = rmul(R0, pe) + ld3(qq);
            const double ex = s2 * w.x, ey = s2 * w.y, ez = s2 * w.z, ew = c2;
            const double x0 = qq[3], y0 = qq[4], z0 = qq[5], w0 = qq[6];
            double rx = w0 * ex + x0 * ew + y0 * ez - z0 * ey;
            double ry = w0 * ey - x0 * ez + y0 * ew + z0 * ex;
            double rz = w0 * ez + x0 * ey - y0 * ex + z0 * ew;
            double rw = w0 * ew - x0 * ex - y0 * ey - z0 * ez;
            if (rx * x0 + ry * y0 + rz * z0 + rw * w0 < 0) {
                rx = -rx, ry = -ry, rz = -rz, rw = -rw;
            }
            const double nn = sqrt(rx * rx + ry * ry + rz * rz + rw * rw);
            st3(qn + iqx, pn);
            qn[iqx + 3] = rx / nn;
            qn[iqx + 4] = ry / nn;
            qn[iqx + 5] = rz / nn;
            qn[iqx + 6] = rw / nn;
        }
        Grp<G>::sync();
        if (io.v_next)
            for (int k = tid; k < nv; k += G)
                io.v_next[(size_t)inst * nv + k] = sm.d[k];
        if (io.q_next)
            for (int k = tid; k < nq; k += G)
                io.q_next[(size_t)inst * nq + k] = qn[k];
        if (io.q_cmd) {
            double* o = io.q_cmd + (size_t)inst * nv;
            if (z.P->floating_base) {
                if (tid == 0) {
                    const double vx = qn[3], vy = qn[4], vz = qn[5], ww = qn[6];
                    double nrm = sqrt(vx * vx + vy * vy + vz * vz);
                    double ax = 1, ay = 0, az = 0, ang = 0;
                    if (nrm != 0) {
                        ang = 2 * atan2(nrm, fabs(ww));
                        if (ww < 0)
                            nrm = -nrm;
                        ax = vx / nrm, ay = vy / nrm, az = vz / nrm;
                    }
                    o[0] = qn[0], o[1] = qn[1], o[2] = qn[2];
                    o[3] = ang * ax, o[4] = ang * ay, o[5] = ang * az;
                }
                for (int k = tid; k < na; k += G)
                    o[6 + k] = qn[7 + k];
            }
            else
                for (int k = tid; k < nv; k += G)
                    o[k] = qn[k];
        }
        if (io.momentum)
            for (int k = tid; k < 3; k += G)
                io.momentum[(size_t)inst * 3 + k] = sm.sc[34 + k]; // Ag.bottomRows(3) * v = angular part of hg
        if (tid == 0) {
            if (io.status) {
                int st = status;
                if (io.sticky_status) {
                    const int old = io.status[inst];
                    if (old != IWBC_STATUS_OPTIMAL)
                        st = old;
                }
                io.status[inst] = st;
            }
            if (io.iters)
                io.iters[inst] = iters;
            if (io.active) {
                unsigned* a = io.active + (size_t)inst * IWBC_ACTIVE_WORDS;
                unsigned words[IWBC_ACTIVE_WORDS];
                for (int k = 0; k < IWBC_ACTIVE_WORDS; ++k)
                    words[k] = 0;
                for (int k = z.S->me; k < iq; ++k) {
                    const int r = sm.act[k];
                    if (r >= 0 && r < 32 * IWBC_ACTIVE_WORDS)
                        words[r >> 5] |= 1u << (r & 31);
                }
                for (int k = 0; k < IWBC_ACTIVE_WORDS; ++k)
                    a[k] = words[k];
            }
        }
        if (io.q_rw && status == IWBC_STATUS_OPTIMAL) { // open-loop roll-out: the integrated state becomes the next tick's state (a failed QP freezes the instance)
            double* qn2 = sm.x_old;
            for (int k = tid; k < nq; k += G)
                io.q_rw[(size_t)inst * nq + k] = qn2[k];
            for (int k = tid; k < nv; k += G)
                io.v_rw[(size_t)inst * nv + k] = sm.d[k];
        }
    }

    // =================================================================================== one instance
    // load -> k_dyn -> k_asm -> k_qp -> k_dec for instance `inst`, executed by one group of G threads
    template <int G>
    __device__ __forceinline__ void cycle_instance(const iwbc_problem_t* __restrict__ gP, const Sm& sm, const Sizes& z, const IO& io, int inst, int tid, int& flip)
    {
        const int nv = z.nv, nq = z.nq, n = z.n, na = z.P->na, nu = nv - na;
        // ---- load state and references (coalesced: the instance's q / v / refs are contiguous)
        const double* gq = io.q + (size_t)inst * nq;
        const double* gv = io.v + (size_t)inst * nv;
        const double* gr = io.refs + (size_t)inst * io.refs_stride;
        for (int k = tid; k < nq; k += G)
            sm.qs[k] = gq[k];
        for (int k = tid; k < nv; k += G)
            sm.vs[k] = gv[k];
        for (int k = tid; k < z.nref; k += G)
            sm.ref[k] = gr[k];
        Grp<G>::sync();
        long long stamps[8];
        long long* cyc = (io.dbg_cycles && inst == 0) ? io.dbg_cycles : (io.prof ? stamps : nullptr);
        if (cyc && tid == 0)
            cyc[0] = clock64();

        k_dyn<G>(gP, sm, z, tid, flip);
        if (cyc && tid == 0)
            cyc[1] = clock64();
        k_asm<G>(gP, sm, z, tid, flip);
        if (cyc && tid == 0)
            cyc[2] = clock64();

        if (io.dbg_M && inst == 0) {
            for (int e = tid; e < nv * nv; e += G)
                io.dbg_M[e] = sm.M[(e / nv) * z.ldM + (e % nv)];
            for (int k = tid; k < nv; k += G)
                io.dbg_nle[k] = sm.nle[k];
            for (int e = tid; e < n * n; e += G) {
                const int i = e / n, j = e % n;
                io.dbg_H[e] = i >= j ? sm.J[i + j * z.ldJ] : sm.J[j + i * z.ldJ];
            }
            for (int k = tid; k < n; k += G)
                io.dbg_g[k] = sm.g[k];
            if (io.dbg_A) {
                for (int e = tid; e < z.rows_dense * nv; e += G)
                    io.dbg_A[e] = sm.A[e];
                for (int k = tid; k < z.rows_dense; k += G)
                    io.dbg_b[k] = sm.b[k];
            }
            Grp<G>::sync();
        }

        int iters = 0, iq = 0;
        const int status = gi_solve<G, FusedRows>(gP, sm, z, (const FusedRows*)nullptr, z.S->me, z.S->mi, tid, flip, iters, iq, cyc);
        if (cyc && tid == 0)
            cyc[5] = clock64();

        k_dec<G>(gP, sm, z, io, inst, tid, status, iters, iq, (const double*)nullptr);
        if (cyc && tid == 0) {
            cyc[6] = clock64();
#ifndef IWBC_CPU_SIM
            if (io.prof) {
                for (int k = 0; k < 6; ++k)
                    atomicAdd(io.prof + k, (unsigned long long)(cyc[k + 1] - cyc[k]));
                atomicAdd(io.prof + 6, 1ull);
            }
#endif
        }
    }

} // namespace dev
} // namespace iwbc
