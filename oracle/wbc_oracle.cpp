// wbc_oracle.cpp -- CPU restatement of one inria_wbc control cycle.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle and the CPU baseline (`bench.py` cpu_baseline / --impl reference).
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it; the product
// library (inria_wbc_b200/csrc) never links, includes or calls anything in oracle/.
//
// PARITY UNPINNED at the third-party boundary: the arithmetic of this path lives in pinocchio, tsid
// and eiquadprog, none of which is vendored, pinned or installable here; this
// file restates their published algorithms in the reference's own order of operations (SURVEY.md
// Appendix A) with scalar FP64 loops and dense "Eigen-shaped" matrices.  It is pinned only by
// (1) the reference's Franka known-answer value (tests/ref_test_franka.yaml:13-15, checked in
// tests/test_franka_kat.py) and (2) algorithm-independent identities (tests/test_oracle_*.py).
//
// What follows which reference lines:
//   cycle order, decode, integrate, repack ....... src/controllers/controller.cpp:231-313
//   SE3 task ...................................... example_project/src/tsid/ex_task.cpp:175-247
//   momentum task ................................. src/tsid/task-momentum-equality.cpp:144-174
//   self-collision task ........................... src/tsid/task-self-collision.cpp:84-204
//   which pinocchio algorithms, in which order .... src/utils/robot_model.cpp:83-113
//   everything upstream (computeAllTerms, Contact6d, formulation, SolverHQuadProgFast,
//   EiquadprogFast::solve_quadprog) ............... SURVEY.md Appendix A.1-A.7
//
// Deliberately *different in structure* from the CUDA path (local-frame recursive dynamics, dense
// HQPData, dense CI, full R) so that agreement between the two is evidence, not tautology.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/iwbc_problem.h"

typedef double real_t; // the problem description (iwbc_problem_t) always holds plain doubles

#ifdef ORACLE_COUNT_FLOPS
// instrumented build (make -C oracle count): every `double` below counts its own arithmetic per cycle phase (count_scalar.hpp)
#include "count_scalar.hpp"
namespace flopcount {
thread_local long counts[NPH][NOP];
thread_local int phase = 0;
}
#define double cdouble
#define ORACLE_PHASE(p) (flopcount::phase = flopcount::p)
#else
#define ORACLE_PHASE(p) ((void)0)
#endif

namespace {

typedef std::vector<double> Vec;

// instrumentation of the active-set loop (oracle_qp_stats): solves, l1 passes, step directions (l2a), dropped constraints
std::atomic<long> g_stat[4];

// ---------------------------------------------------------------------------------- small math
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct M3 {
    double m[9]; // row-major
};
inline V3 operator*(const M3& R, V3 v)
{
    return {R.m[0] * v.x + R.m[1] * v.y + R.m[2] * v.z, R.m[3] * v.x + R.m[4] * v.y + R.m[5] * v.z,
        R.m[6] * v.x + R.m[7] * v.y + R.m[8] * v.z};
}
inline V3 tmul(const M3& R, V3 v) // R^T v
{
    return {R.m[0] * v.x + R.m[3] * v.y + R.m[6] * v.z, R.m[1] * v.x + R.m[4] * v.y + R.m[7] * v.z,
        R.m[2] * v.x + R.m[5] * v.y + R.m[8] * v.z};
}
inline M3 operator*(const M3& A, const M3& B)
{
    M3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            C.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
    return C;
}
inline M3 transpose(const M3& A)
{
    M3 T;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            T.m[3 * i + j] = A.m[3 * j + i];
    return T;
}
inline M3 identity3() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }

struct SE3 {
    M3 R;
    V3 p;
};
inline SE3 operator*(const SE3& a, const SE3& b) { return {a.R * b.R, a.p + a.R * b.p}; }
inline SE3 inverse(const SE3& a)
{
    M3 Rt = transpose(a.R);
    return {Rt, -1.0 * (Rt * a.p)};
}

struct Motion {
    V3 lin, ang;
}; // pinocchio order: linear, angular
struct Force {
    V3 lin, ang;
};
inline Motion operator+(Motion a, Motion b) { return {a.lin + b.lin, a.ang + b.ang}; }
inline Force operator+(Force a, Force b) { return {a.lin + b.lin, a.ang + b.ang}; }
inline Motion act(const SE3& X, Motion m)
{
    V3 w = X.R * m.ang;
    return {X.R * m.lin + cross(X.p, w), w};
}
inline Motion actInv(const SE3& X, Motion m) { return {tmul(X.R, m.lin - cross(X.p, m.ang)), tmul(X.R, m.ang)}; }
inline Force act(const SE3& X, Force f)
{
    V3 l = X.R * f.lin;
    return {l, X.R * f.ang + cross(X.p, l)};
}
inline Motion mcross(Motion a, Motion b) { return {cross(a.lin, b.ang) + cross(a.ang, b.lin), cross(a.ang, b.ang)}; }
inline Force mcrossf(Motion v, Force f) { return {cross(v.ang, f.lin), cross(v.ang, f.ang) + cross(v.lin, f.lin)}; }

struct Inertia {
    double m;
    V3 c;
    M3 I; // about c
};
inline Force operator*(const Inertia& Y, Motion v)
{
    V3 l = Y.m * (v.lin - cross(Y.c, v.ang));
    return {l, Y.I * v.ang + cross(Y.c, l)};
}
inline Inertia act(const SE3& X, const Inertia& Y) { return {Y.m, X.R * Y.c + X.p, X.R * Y.I * transpose(X.R)}; }
inline Inertia operator+(const Inertia& a, const Inertia& b)
{
    if (a.m == 0)
        return b;
    if (b.m == 0)
        return a;
    Inertia o;
    o.m = a.m + b.m;
    o.c = (1.0 / o.m) * (a.m * a.c + b.m * b.c);
    for (int k = 0; k < 9; ++k)
        o.I.m[k] = 0;
    const Inertia* both[2] = {&a, &b};
    for (const Inertia* x : both) {
        V3 d = x->c - o.c;
        double dd = dot(d, d), dv[3] = {d.x, d.y, d.z};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                o.I.m[3 * i + j] += x->I.m[3 * i + j] + x->m * ((i == j ? dd : double(0.0)) - dv[i] * dv[j]);
    }
    return o;
}

inline M3 exp3(V3 w)
{
    double t2 = dot(w, w), t = std::sqrt(t2);
    double a, b; // R = I + a [w] + b [w]^2
    if (t < 1.220703125e-4) {
        a = 1 - t2 / 6;
        b = 0.5 - t2 / 24;
    }
    else {
        a = std::sin(t) / t;
        b = (1 - std::cos(t)) / t2;
    }
    M3 W = {{0, -w.z, w.y, w.z, 0, -w.x, -w.y, w.x, 0}};
    M3 W2 = W * W;
    M3 R = identity3();
    for (int k = 0; k < 9; ++k)
        R.m[k] += a * W.m[k] + b * W2.m[k];
    return R;
}
// log3 as pinocchio: theta from the trace, axis from the antisymmetric part (with the pi branch)
inline V3 log3(const M3& R)
{
    double tr = R.m[0] + R.m[4] + R.m[8];
    double c = std::min(double(1.0), std::max(double(-1.0), 0.5 * (tr - 1)));
    double theta = std::acos(c);
    V3 v = {R.m[7] - R.m[5], R.m[2] - R.m[6], R.m[3] - R.m[1]};
    if (theta < 1.220703125e-4) // TaylorSeriesExpansion<double>::precision<3>() = eps^(1/4)
        return 0.5 * v;
    if (theta > M_PI - 1e-2) {
        // near pi: use the diagonal
        double cphi = -(tr - 1) / 2; // = -cos(theta)
        double beta = theta * theta / (1 + cphi);
        V3 t = {(R.m[0] + cphi) * beta, (R.m[4] + cphi) * beta, (R.m[8] + cphi) * beta};
        V3 r;
        r.x = (v.x > 0 ? 1 : -1) * (t.x > 0 ? std::sqrt(t.x) : double(0));
        r.y = (v.y > 0 ? 1 : -1) * (t.y > 0 ? std::sqrt(t.y) : double(0));
        r.z = (v.z > 0 ? 1 : -1) * (t.z > 0 ? std::sqrt(t.z) : double(0));
        return r;
    }
    return (theta / (2 * std::sin(theta))) * v;
}
// log6: twist (lin, ang) with M = exp6(twist)
inline Motion log6(const SE3& M)
{
    V3 w = log3(M.R);
    double t2 = dot(w, w), t = std::sqrt(t2);
    double alpha, beta;
    if (t < 1.220703125e-4) {
        alpha = 1 - t2 / 12 - t2 * t2 / 720;
        beta = 1.0 / 12 + t2 / 720;
    }
    else {
        double st = std::sin(t), ct = std::cos(t);
        alpha = t * st / (2 * (1 - ct));
        beta = 1 / t2 - st / (2 * t * (1 - ct));
    }
    // v = alpha p - 0.5 w x p + beta (w.p) w
    V3 lin = alpha * M.p - 0.5 * cross(w, M.p) + (beta * dot(w, M.p)) * w;
    return {lin, w};
}
inline SE3 exp6(Motion nu)
{
    V3 w = nu.ang, v = nu.lin;
    double t2 = dot(w, w), t = std::sqrt(t2);
    double a, b, c; // R = I + a W + b W^2 ; p = (I + b W + c W^2) v
    if (t < 1.220703125e-4) {
        a = 1 - t2 / 6;
        b = 0.5 - t2 / 24;
        c = 1.0 / 6 - t2 / 120;
    }
    else {
        a = std::sin(t) / t;
        b = (1 - std::cos(t)) / t2;
        c = (t - std::sin(t)) / (t2 * t);
    }
    SE3 M;
    M.R = exp3(w);
    (void)a;
    V3 wv = cross(w, v);
    M.p = v + b * wv + c * cross(w, wv);
    return M;
}

// ------------------------------------------------------------------------------ dense helpers
struct Mat { // row-major
    int r = 0, c = 0;
    Vec a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

// ------------------------------------------------------------------- eiquadprog-fast restated
// SURVEY.md Appendix A.7.  min 0.5 x'Hx + g'x  s.t. CE x + ce0 = 0, CI x + ci0 >= 0.
enum { QP_OPTIMAL = 0, QP_INFEASIBLE = 1, QP_UNBOUNDED = 2, QP_MAX_ITER = 3, QP_REDUNDANT_EQ = 4 };

struct QPWork {
    int n = 0;
    Mat J, R;
    Vec d, z, r, u, s, np, x_old, u_old;
    std::vector<int> A, A_old, iai, iaexcl;
    int iq = 0, iter = 0;
    double f_value = 0;
};

inline double gi_distance(double a, double b)
{
    double a1 = std::fabs(a), b1 = std::fabs(b);
    if (a1 > b1) {
        double t = b1 / a1;
        return a1 * std::sqrt(1.0 + t * t);
    }
    else if (b1 > a1) {
        double t = a1 / b1;
        return b1 * std::sqrt(1.0 + t * t);
    }
    return a1 * std::sqrt(2.0);
}

bool gi_add_constraint(QPWork& w, double& R_norm)
{
    const int n = w.n;
    for (int j = n - 1; j >= w.iq + 1; --j) {
        double cc = w.d[j - 1], ss = w.d[j];
        double h = gi_distance(cc, ss);
        if (h == 0.0)
            continue;
        w.d[j] = 0.0;
        ss = ss / h;
        cc = cc / h;
        if (cc < 0.0) {
            cc = -cc;
            ss = -ss;
            w.d[j - 1] = -h;
        }
        else
            w.d[j - 1] = h;
        double xny = ss / (1.0 + cc);
        for (int k = 0; k < n; ++k) {
            double t1 = w.J(k, j - 1), t2 = w.J(k, j);
            w.J(k, j - 1) = t1 * cc + t2 * ss;
            w.J(k, j) = xny * (t1 + w.J(k, j - 1)) - t2;
        }
    }
    w.iq++;
    for (int i = 0; i < w.iq; ++i)
        w.R(i, w.iq - 1) = w.d[i];
    if (std::fabs(w.d[w.iq - 1]) <= std::numeric_limits<double>::epsilon() * R_norm)
        return false;
    R_norm = std::max(R_norm, std::fabs(w.d[w.iq - 1]));
    return true;
}

void gi_delete_constraint(QPWork& w, int me, int l)
{
    const int n = w.n;
    int qq = 0;
    for (int i = me; i < w.iq; ++i)
        if (w.A[i] == l) {
            qq = i;
            break;
        }
    for (int i = qq; i < w.iq - 1; ++i) {
        w.A[i] = w.A[i + 1];
        w.u[i] = w.u[i + 1];
        for (int k = 0; k < n; ++k)
            w.R(k, i) = w.R(k, i + 1);
    }
    w.A[w.iq - 1] = w.A[w.iq];
    w.u[w.iq - 1] = w.u[w.iq];
    w.A[w.iq] = 0;
    w.u[w.iq] = 0.0;
    for (int j = 0; j < w.iq; ++j)
        w.R(j, w.iq - 1) = 0.0;
    w.iq--;
    if (w.iq == 0)
        return;
    for (int j = qq; j < w.iq; ++j) {
        double cc = w.R(j, j), ss = w.R(j + 1, j);
        double h = gi_distance(cc, ss);
        if (h == 0.0)
            continue;
        cc = cc / h;
        ss = ss / h;
        w.R(j + 1, j) = 0.0;
        if (cc < 0.0) {
            w.R(j, j) = -h;
            cc = -cc;
            ss = -ss;
        }
        else
            w.R(j, j) = h;
        double xny = ss / (1.0 + cc);
        for (int k = j + 1; k < w.iq; ++k) {
            double t1 = w.R(j, k), t2 = w.R(j + 1, k);
            w.R(j, k) = t1 * cc + t2 * ss;
            w.R(j + 1, k) = xny * (t1 + w.R(j, k)) - t2;
        }
        for (int k = 0; k < n; ++k) {
            double t1 = w.J(k, j), t2 = w.J(k, j + 1);
            w.J(k, j) = t1 * cc + t2 * ss;
            w.J(k, j + 1) = xny * (w.J(k, j) + t1) - t2;
        }
    }
}

int solve_quadprog(const Mat& Hin, const Vec& g0, const Mat& CE, const Vec& ce0, const Mat& CI, const Vec& ci0,
    Vec& x, QPWork& w, int max_iter = 1000)
{
    const int n = (int)g0.size(), me = (int)ce0.size(), mi = (int)ci0.size();
    const double inf = std::numeric_limits<double>::infinity();
    const double eps = std::numeric_limits<double>::epsilon();
    w.n = n;
    w.J = Mat(n, n);
    w.R = Mat(n, n);
    w.d.assign(n, 0);
    w.z.assign(n, 0);
    w.r.assign(n, 0);
    w.np.assign(n, 0);
    w.u.assign(me + mi + 1, 0);
    w.u_old.assign(me + mi + 1, 0);
    w.x_old.assign(n, 0);
    w.s.assign(mi, 0);
    w.A.assign(me + mi + 1, 0);
    w.A_old.assign(me + mi + 1, 0);
    w.iai.assign(mi, 0);
    w.iaexcl.assign(mi, 1);
    w.iter = 0;
    x.assign(n, 0.0);

    double c1 = 0;
    for (int i = 0; i < n; ++i)
        c1 += Hin(i, i);
    // LLT
    Mat L = Hin;
    for (int j = 0; j < n; ++j) {
        double sum = L(j, j);
        for (int k = 0; k < j; ++k)
            sum -= L(j, k) * L(j, k);
        if (sum <= 0)
            return QP_UNBOUNDED; // not PD (eiquadprog would produce NaNs)
        double ljj = std::sqrt(sum);
        L(j, j) = ljj;
        for (int i = j + 1; i < n; ++i) {
            double s = L(i, j);
            for (int k = 0; k < j; ++k)
                s -= L(i, k) * L(j, k);
            L(i, j) = s / ljj;
        }
    }
    double R_norm = 1.0;
    // J = L^-T : solve L^T J = I (back substitution per column)
    for (int col = 0; col < n; ++col) {
        for (int i = n - 1; i >= 0; --i) {
            double s = (i == col) ? 1.0 : 0.0;
            for (int k = i + 1; k < n; ++k)
                s -= L(k, i) * w.J(k, col);
            w.J(i, col) = s / L(i, i);
        }
    }
    double c2 = 0;
    for (int i = 0; i < n; ++i)
        c2 += w.J(i, i);
    // x = -H^-1 g0
    {
        Vec y(n);
        for (int i = 0; i < n; ++i) {
            double s = g0[i];
            for (int k = 0; k < i; ++k)
                s -= L(i, k) * y[k];
            y[i] = s / L(i, i);
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int k = i + 1; k < n; ++k)
                s -= L(k, i) * x[k];
            x[i] = s / L(i, i);
        }
        for (int i = 0; i < n; ++i)
            x[i] = -x[i];
    }
    w.f_value = 0;
    for (int i = 0; i < n; ++i)
        w.f_value += 0.5 * g0[i] * x[i];

    auto compute_d = [&]() {
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int i = 0; i < n; ++i)
                s += w.J(i, j) * w.np[i];
            w.d[j] = s;
        }
    };
    auto update_z = [&]() {
        for (int i = 0; i < n; ++i) {
            double s = 0;
            for (int j = w.iq; j < n; ++j)
                s += w.J(i, j) * w.d[j];
            w.z[i] = s;
        }
    };
    auto update_r = [&]() {
        for (int i = w.iq - 1; i >= 0; --i) {
            double s = w.d[i];
            for (int j = i + 1; j < w.iq; ++j)
                s -= w.R(i, j) * w.r[j];
            w.r[i] = s / w.R(i, i);
        }
    };
    auto dotn = [&](const Vec& a, const Vec& b) {
        double s = 0;
        for (int i = 0; i < n; ++i)
            s += a[i] * b[i];
        return s;
    };

    int& iq = w.iq;
    iq = 0;
    for (int i = 0; i < me; ++i) {
        for (int k = 0; k < n; ++k)
            w.np[k] = CE(i, k);
        compute_d();
        update_z();
        update_r();
        double t2 = 0.0;
        if (std::fabs(dotn(w.z, w.z)) > eps)
            t2 = (-dotn(w.np, x) - ce0[i]) / dotn(w.z, w.np);
        for (int k = 0; k < n; ++k)
            x[k] += t2 * w.z[k];
        w.u[iq] = t2;
        for (int k = 0; k < iq; ++k)
            w.u[k] -= t2 * w.r[k];
        w.f_value += 0.5 * (t2 * t2) * dotn(w.z, w.np);
        w.A[i] = -i - 1;
        if (!gi_add_constraint(w, R_norm))
            return QP_REDUNDANT_EQ;
    }
    for (int i = 0; i < mi; ++i)
        w.iai[i] = i;

    int ip = 0, l = 0;
    double ss, psi, t, t1, t2;
l1:
    g_stat[1]++;
    w.iter++;
    if (w.iter >= max_iter)
        return QP_MAX_ITER;
    for (int i = me; i < iq; ++i) {
        ip = w.A[i];
        w.iai[ip] = -1;
    }
    ss = 0.0;
    psi = 0.0;
    ip = 0;
    for (int i = 0; i < mi; ++i) {
        double sum = ci0[i];
        for (int k = 0; k < n; ++k)
            sum += CI(i, k) * x[k];
        w.s[i] = sum;
        w.iaexcl[i] = 1;
        psi += std::min(double(0.0), sum);
    }
    if (std::fabs(psi) <= mi * eps * c1 * c2 * 100.0)
        return QP_OPTIMAL;
    for (int i = 0; i < iq; ++i) {
        w.u_old[i] = w.u[i];
        w.A_old[i] = w.A[i];
    }
    w.x_old = x;
l2:
    for (int i = 0; i < mi; ++i) {
        if (w.s[i] < ss && w.iai[i] != -1 && w.iaexcl[i]) {
            ss = w.s[i];
            ip = i;
        }
    }
    if (ss >= 0.0)
        return QP_OPTIMAL;
    for (int k = 0; k < n; ++k)
        w.np[k] = CI(ip, k);
    w.u[iq] = 0.0;
    w.A[iq] = ip;
l2a:
    g_stat[2]++;
    compute_d();
    update_z();
    update_r();
    l = 0;
    t1 = inf;
    for (int k = me; k < iq; ++k) {
        double tmp;
        if (w.r[k] > 0.0 && ((tmp = w.u[k] / w.r[k]) < t1)) {
            t1 = tmp;
            l = w.A[k];
        }
    }
    if (std::fabs(dotn(w.z, w.z)) > eps)
        t2 = -w.s[ip] / dotn(w.z, w.np);
    else
        t2 = inf;
    t = std::min(t1, t2);
    if (t >= inf)
        return QP_UNBOUNDED; // primal infeasible
    if (t2 >= inf) {
        for (int k = 0; k < iq; ++k)
            w.u[k] -= t * w.r[k];
        w.u[iq] += t;
        w.iai[l] = l;
        g_stat[3]++;
        gi_delete_constraint(w, me, l);
        goto l2a;
    }
    for (int k = 0; k < n; ++k)
        x[k] += t * w.z[k];
    w.f_value += t * dotn(w.z, w.np) * (0.5 * t + w.u[iq]);
    for (int k = 0; k < iq; ++k)
        w.u[k] -= t * w.r[k];
    w.u[iq] += t;
    if (t == t2) {
        if (!gi_add_constraint(w, R_norm)) {
            w.iaexcl[ip] = 0;
            gi_delete_constraint(w, me, ip);
            for (int i = 0; i < mi; ++i)
                w.iai[i] = i;
            for (int i = 0; i < iq; ++i) {
                w.A[i] = w.A_old[i];
                w.iai[w.A[i]] = -1;
                w.u[i] = w.u_old[i];
            }
            x = w.x_old;
            goto l2;
        }
        else
            w.iai[ip] = -1;
        goto l1;
    }
    w.iai[l] = l;
    g_stat[3]++;
    gi_delete_constraint(w, me, l);
    {
        double sum = ci0[ip];
        for (int k = 0; k < n; ++k)
            sum += CI(ip, k) * x[k];
        w.s[ip] = sum;
    }
    goto l2a;
}

// ------------------------------------------------------------------------------- robot dynamics
struct Dyn {
    int nj, nv;
    std::vector<SE3> liMi, oMi;
    std::vector<Motion> v, a_gf, a; // joint-frame spatial velocity, accel with gravity field, drift accel (no gravity)
    std::vector<Inertia> Ycrb;
    std::vector<Force> f;
    Mat M, J, Jcom, Ag; // M nv x nv ; J 6 x nv world ; Jcom 3 x nv ; Ag 6 x nv (lin rows 0-2, ang 3-5)
    Vec nle;
    V3 com, vcom, acom;
    Force hg, dhg; // centroidal momentum and its drift (zero joint acceleration)
    std::vector<Force> Fcol; // per dof: Ycrb * S in the joint frame (before propagation)
};

template <class Real>
inline SE3 load_se3(const Real* p12)
{
    SE3 X;
    for (int i = 0; i < 9; ++i)
        X.R.m[i] = p12[i];
    X.p = {p12[9], p12[10], p12[11]};
    return X;
}
inline M3 quat_to_R(double x, double y, double z, double w)
{
    return {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
        2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
}
inline M3 axis_angle(V3 a, double th)
{
    double c = std::cos(th), s = std::sin(th), v = 1 - c;
    return {{a.x * a.x * v + c, a.x * a.y * v - a.z * s, a.x * a.z * v + a.y * s, a.x * a.y * v + a.z * s,
        a.y * a.y * v + c, a.y * a.z * v - a.x * s, a.x * a.z * v - a.y * s, a.y * a.z * v + a.x * s, a.z * a.z * v + c}};
}
inline Inertia joint_inertia(const iwbc_problem_t& P, int j)
{
    Inertia Y;
    Y.m = P.mass[j];
    Y.c = {P.com[j][0], P.com[j][1], P.com[j][2]};
    const real_t* I = P.inertia[j];
    Y.I = {{I[0], I[1], I[2], I[1], I[3], I[4], I[2], I[4], I[5]}};
    return Y;
}
// motion subspace column k (0..nvj-1) of joint j in the joint frame
inline Motion subspace(const iwbc_problem_t& P, int j, int k)
{
    Motion S = {{0, 0, 0}, {0, 0, 0}};
    V3 ax = {P.axis[j][0], P.axis[j][1], P.axis[j][2]};
    if (P.jtype[j] == IWBC_J_REVOLUTE)
        S.ang = ax;
    else if (P.jtype[j] == IWBC_J_PRISMATIC)
        S.lin = ax;
    else {
        double e[6] = {0, 0, 0, 0, 0, 0};
        e[k] = 1;
        S.lin = {e[0], e[1], e[2]};
        S.ang = {e[3], e[4], e[5]};
    }
    return S;
}
inline int joint_nv(const iwbc_problem_t& P, int j) { return P.jtype[j] == IWBC_J_FREEFLYER ? 6 : 1; }

// RobotWrapper::computeAllTerms (Appendix A.1): FK, CRBA, nle, Jacobians, CoM (0,1,2 order), drift, ccrba
void compute_all_terms(const iwbc_problem_t& P, const double* q, const double* vq, Dyn& D)
{
    const int nj = P.nj, nv = P.nv;
    D.nj = nj;
    D.nv = nv;
    D.liMi.resize(nj);
    D.oMi.resize(nj);
    D.v.resize(nj);
    D.a_gf.resize(nj);
    D.a.resize(nj);
    D.Ycrb.resize(nj);
    D.f.resize(nj);
    D.Fcol.resize(nv);
    D.M = Mat(nv, nv);
    D.J = Mat(6, nv);
    D.Jcom = Mat(3, nv);
    D.Ag = Mat(6, nv);
    D.nle.assign(nv, 0.0);
    const Motion zero = {{0, 0, 0}, {0, 0, 0}};
    const Motion a_gf0 = {{-P.gravity[0], -P.gravity[1], -P.gravity[2]}, {0, 0, 0}};

    // ---- forward pass
    for (int j = 0; j < nj; ++j) {
        SE3 jM;
        jM.R = identity3();
        jM.p = {0, 0, 0};
        V3 ax = {P.axis[j][0], P.axis[j][1], P.axis[j][2]};
        Motion vJ = zero;
        if (P.jtype[j] == IWBC_J_FREEFLYER) {
            const double* qq = q + P.idx_q[j];
            jM.R = quat_to_R(qq[3], qq[4], qq[5], qq[6]);
            jM.p = {qq[0], qq[1], qq[2]};
            const double* vv = vq + P.idx_v[j];
            vJ = {{vv[0], vv[1], vv[2]}, {vv[3], vv[4], vv[5]}};
        }
        else if (P.jtype[j] == IWBC_J_REVOLUTE) {
            jM.R = axis_angle(ax, q[P.idx_q[j]]);
            vJ.ang = vq[P.idx_v[j]] * ax;
        }
        else {
            jM.p = q[P.idx_q[j]] * ax;
            vJ.lin = vq[P.idx_v[j]] * ax;
        }
        D.liMi[j] = load_se3(P.jplace[j]) * jM;
        int p = P.parent[j];
        D.oMi[j] = p >= 0 ? D.oMi[p] * D.liMi[j] : D.liMi[j];
        Motion vp = p >= 0 ? actInv(D.liMi[j], D.v[p]) : zero;
        D.v[j] = vp + vJ;
        Motion bias = mcross(D.v[j], vJ);
        D.a_gf[j] = actInv(D.liMi[j], p >= 0 ? D.a_gf[p] : a_gf0) + bias;
        D.a[j] = (p >= 0 ? actInv(D.liMi[j], D.a[p]) : zero) + bias;
        Inertia Y = joint_inertia(P, j);
        D.Ycrb[j] = Y;
        D.f[j] = Y * D.a_gf[j] + mcrossf(D.v[j], Y * D.v[j]);
        // world joint Jacobian columns
        for (int k = 0; k < joint_nv(P, j); ++k) {
            Motion S = act(D.oMi[j], subspace(P, j, k));
            int c = P.idx_v[j] + k;
            D.J(0, c) = S.lin.x;
            D.J(1, c) = S.lin.y;
            D.J(2, c) = S.lin.z;
            D.J(3, c) = S.ang.x;
            D.J(4, c) = S.ang.y;
            D.J(5, c) = S.ang.z;
        }
    }
    // ---- CoM position / velocity / drift acceleration (centerOfMass(q, v, 0))
    {
        double mtot = 0;
        V3 c = {0, 0, 0}, vc = {0, 0, 0}, ac = {0, 0, 0};
        for (int j = 0; j < nj; ++j) {
            Inertia Y = joint_inertia(P, j);
            mtot += Y.m;
            c = c + Y.m * (D.oMi[j].R * Y.c + D.oMi[j].p);
            V3 vloc = D.v[j].lin + cross(D.v[j].ang, Y.c);
            vc = vc + Y.m * (D.oMi[j].R * vloc);
            V3 aloc = D.a[j].lin + cross(D.a[j].ang, Y.c) + cross(D.v[j].ang, vloc);
            ac = ac + Y.m * (D.oMi[j].R * aloc);
        }
        D.com = (1.0 / mtot) * c;
        D.vcom = (1.0 / mtot) * vc;
        D.acom = (1.0 / mtot) * ac;
    }
    // ---- backward pass: nle (RNEA) and CRBA with force-set propagation
    {
        std::vector<Force> f = D.f;
        for (int j = nj - 1; j >= 0; --j) {
            for (int k = 0; k < joint_nv(P, j); ++k) {
                Motion S = subspace(P, j, k);
                D.nle[P.idx_v[j] + k] = dot(S.lin, f[j].lin) + dot(S.ang, f[j].ang);
            }
            if (P.parent[j] >= 0)
                f[P.parent[j]] = f[P.parent[j]] + act(D.liMi[j], f[j]);
        }
    }
    {
        // Fset[c] = force column c, expressed in the frame of the joint currently being processed
        std::vector<Force> Fset(nv);
        for (int j = nj - 1; j >= 0; --j) {
            const int c0 = P.idx_v[j], nvj = joint_nv(P, j);
            const int clast = P.idx_v[P.last_desc[j]] + joint_nv(P, P.last_desc[j]); // one past the subtree's last dof
            for (int k = 0; k < nvj; ++k) {
                Fset[c0 + k] = D.Ycrb[j] * subspace(P, j, k);
                D.Fcol[c0 + k] = Fset[c0 + k];
            }
            for (int k = 0; k < nvj; ++k) {
                Motion S = subspace(P, j, k);
                for (int c = c0; c < clast; ++c)
                    D.M(c0 + k, c) = dot(S.lin, Fset[c].lin) + dot(S.ang, Fset[c].ang);
            }
            int p = P.parent[j];
            if (p >= 0) {
                D.Ycrb[p] = D.Ycrb[p] + act(D.liMi[j], D.Ycrb[j]);
                for (int c = c0; c < clast; ++c)
                    Fset[c] = act(D.liMi[j], Fset[c]);
            }
        }
        // symmetrise (RobotWrapper::computeAllTerms; src/utils/robot_model.cpp:103)
        for (int i = 0; i < nv; ++i)
            for (int j = 0; j < i; ++j)
                D.M(i, j) = D.M(j, i);
    }
    // ---- Jcom (jacobianCenterOfMass) from subtree masses / CoMs
    {
        std::vector<double> msub(nj);
        std::vector<V3> csub(nj); // mass-weighted world CoM sum
        for (int j = 0; j < nj; ++j) {
            msub[j] = P.mass[j];
            V3 cl = {P.com[j][0], P.com[j][1], P.com[j][2]};
            csub[j] = P.mass[j] * (D.oMi[j].R * cl + D.oMi[j].p);
        }
        for (int j = nj - 1; j >= 0; --j)
            if (P.parent[j] >= 0) {
                msub[P.parent[j]] += msub[j];
                csub[P.parent[j]] = csub[P.parent[j]] + csub[j];
            }
        double mtot = 0;
        for (int j = 0; j < nj; ++j)
            if (P.parent[j] < 0)
                mtot += msub[j];
        for (int j = 0; j < nj; ++j) {
            V3 cs = msub[j] > 0 ? (1.0 / msub[j]) * csub[j] : D.oMi[j].p;
            for (int k = 0; k < joint_nv(P, j); ++k) {
                Motion S = subspace(P, j, k);
                V3 w = D.oMi[j].R * S.ang, l = D.oMi[j].R * S.lin;
                V3 col = (msub[j] / mtot) * (l + cross(w, cs - D.oMi[j].p));
                int c = P.idx_v[j] + k;
                D.Jcom(0, c) = col.x;
                D.Jcom(1, c) = col.y;
                D.Jcom(2, c) = col.z;
            }
        }
    }
    // ---- ccrba: Ag columns = composite momentum of the subtree per unit dof velocity, about the CoM
    for (int j = 0; j < nj; ++j)
        for (int k = 0; k < joint_nv(P, j); ++k) {
            int c = P.idx_v[j] + k;
            Force F = act(D.oMi[j], D.Fcol[c]); // world, about the origin
            V3 ang = F.ang - cross(D.com, F.lin);
            D.Ag(0, c) = F.lin.x;
            D.Ag(1, c) = F.lin.y;
            D.Ag(2, c) = F.lin.z;
            D.Ag(3, c) = ang.x;
            D.Ag(4, c) = ang.y;
            D.Ag(5, c) = ang.z;
        }
    // ---- centroidal momentum and its time variation at zero joint acceleration (drift, no gravity)
    {
        Force h = {{0, 0, 0}, {0, 0, 0}}, dh = h;
        for (int j = 0; j < nj; ++j) {
            Inertia Y = joint_inertia(P, j);
            Force hj = Y * D.v[j];
            Force fj = Y * D.a[j] + mcrossf(D.v[j], hj);
            h = h + act(D.oMi[j], hj);
            dh = dh + act(D.oMi[j], fj);
        }
        D.hg = {h.lin, h.ang - cross(D.com, h.lin)};
        D.dhg = {dh.lin, dh.ang - cross(D.com, dh.lin)};
    }
}

// ancestors test: is joint a an ancestor-or-self of joint j (DFS numbering)
inline bool supports(const iwbc_problem_t& P, int a, int j) { return a <= j && j <= P.last_desc[a]; }

struct FrameKin {
    SE3 oMf;
    Motion v, a_classic; // local frame
    Mat Jlocal;          // 6 x nv
};
void frame_kinematics(const iwbc_problem_t& P, const Dyn& D, int frame, FrameKin& K, bool want_J = true)
{
    const int pj = P.fparent[frame];
    SE3 fM = load_se3(P.fplace[frame]);
    const Motion zero = {{0, 0, 0}, {0, 0, 0}};
    K.oMf = pj >= 0 ? D.oMi[pj] * fM : fM;
    K.v = pj >= 0 ? actInv(fM, D.v[pj]) : zero;
    Motion a = pj >= 0 ? actInv(fM, D.a[pj]) : zero;
    a.lin = a.lin + cross(K.v.ang, K.v.lin); // frameClassicAcceleration
    K.a_classic = a;
    if (!want_J)
        return;
    K.Jlocal = Mat(6, P.nv);
    if (pj < 0)
        return;
    for (int j = 0; j < P.nj; ++j) {
        if (!supports(P, j, pj))
            continue;
        for (int k = 0; k < joint_nv(P, j); ++k) {
            int c = P.idx_v[j] + k;
            Motion col = {{D.J(0, c), D.J(1, c), D.J(2, c)}, {D.J(3, c), D.J(4, c), D.J(5, c)}};
            Motion l = actInv(K.oMf, col);
            K.Jlocal(0, c) = l.lin.x;
            K.Jlocal(1, c) = l.lin.y;
            K.Jlocal(2, c) = l.lin.z;
            K.Jlocal(3, c) = l.ang.x;
            K.Jlocal(4, c) = l.ang.y;
            K.Jlocal(5, c) = l.ang.z;
        }
    }
}

// TaskSE3Equality::compute, local frame (ex_task.cpp:175-247, m_local_frame = true)
void se3_task_rows(const iwbc_problem_t& P, const Dyn& D, int frame, const real_t* kp, const real_t* kd, const double* ref,
    Mat& A6, double* b6)
{
    FrameKin K;
    frame_kinematics(P, D, frame, K);
    SE3 Mref;
    Mref.p = {ref[0], ref[1], ref[2]};
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            Mref.R.m[3 * r + c] = ref[3 + 3 * c + r];
    Motion err = log6(inverse(K.oMf) * Mref); // errorInSE3
    Motion vref = {{ref[12], ref[13], ref[14]}, {ref[15], ref[16], ref[17]}};
    Motion aref = {{ref[18], ref[19], ref[20]}, {ref[21], ref[22], ref[23]}};
    // m_wMl has the frame's rotation and zero translation
    V3 vr_l = tmul(K.oMf.R, vref.lin), vr_a = tmul(K.oMf.R, vref.ang);
    V3 ar_l = tmul(K.oMf.R, aref.lin), ar_a = tmul(K.oMf.R, aref.ang);
    double e[6] = {err.lin.x, err.lin.y, err.lin.z, err.ang.x, err.ang.y, err.ang.z};
    double ve[6] = {vr_l.x - K.v.lin.x, vr_l.y - K.v.lin.y, vr_l.z - K.v.lin.z, vr_a.x - K.v.ang.x, vr_a.y - K.v.ang.y,
        vr_a.z - K.v.ang.z};
    double ar[6] = {ar_l.x, ar_l.y, ar_l.z, ar_a.x, ar_a.y, ar_a.z};
    double dr[6] = {K.a_classic.lin.x, K.a_classic.lin.y, K.a_classic.lin.z, K.a_classic.ang.x, K.a_classic.ang.y,
        K.a_classic.ang.z};
    A6 = K.Jlocal;
    for (int i = 0; i < 6; ++i)
        b6[i] = kp[i] * e[i] + kd[i] * ve[i] + ar[i] - dr[i];
}

// TaskJointPosVelAccBounds::computeAccLimits (Appendix A.5); note m_dt = 2 dt in the upstream constructor
void joint_acc_bounds(const iwbc_problem_t& P, const double* q, const double* v, double* lbv, double* ubv)
{
    const int na = P.na, nq = P.nq, nv = P.nv;
    const double dt = 2.0 * P.dt;
    for (int i = 0; i < na; ++i) {
        const double qi = q[nq - na + i], dq = v[nv - na + i];
        const double qmin = P.q_lo[i], qmax = P.q_hi[i], dqmax = P.v_max[i], ddqmax = P.v_max[i] / P.dt;
        const double m = -dq / dt;
        // position
        double lbP = -1e10, ubP = 1e10;
        double ddqMax_q3 = 2 * (qmax - qi - dt * dq) / (dt * dt);
        double ddqMin_q3 = 2 * (qmin - qi - dt * dq) / (dt * dt);
        if (dq <= 0.0) {
            ubP = ddqMax_q3;
            if (ddqMin_q3 < m)
                lbP = ddqMin_q3;
            else if (qi != qmin)
                lbP = std::max((dq * dq) / (2 * (qi - qmin)), m);
            else
                lbP = 1e6;
        }
        else {
            lbP = ddqMin_q3;
            if (ddqMax_q3 > m)
                ubP = ddqMax_q3;
            else if (qi != qmax)
                ubP = std::min(-(dq * dq) / (2 * (qmax - qi)), m);
            else
                ubP = -1e6;
        }
        // velocity
        double lbV = (-dqmax - dq) / dt, ubV = (dqmax - dq) / dt;
        // viability
        double a2 = 2 * dt * dt;
        double b1 = 2 * dt * dq + ddqmax * dt * dt, b2 = 2 * dt * dq - ddqmax * dt * dt;
        double c1 = dq * dq - 2 * ddqmax * (qmax - (qi + dt * dq));
        double c2 = dq * dq - 2 * ddqmax * ((qi + dt * dq) - qmin);
        double d1 = b1 * b1 - 2 * a2 * c1, d2 = b2 * b2 - 2 * a2 * c2;
        double ddq1 = d1 >= 0 ? (-b1 + std::sqrt(d1)) / a2 : m;
        double ddq2 = d2 >= 0 ? (-b2 - std::sqrt(d2)) / a2 : m;
        double ubVia = std::max(ddq1, m), lbVia = std::min(ddq2, m);
        // acceleration
        double lbA = -ddqmax, ubA = ddqmax;
        double lb = std::max(std::max(lbP, lbVia), std::max(lbV, lbA));
        double ub = std::min(std::min(ubP, ubVia), std::min(ubV, ubA));
        if (ub < lb) { // conflict: position bounds win
            if (ub == ubP)
                lb = ub;
            else
                ub = lb;
        }
        lbv[i] = lb;
        ubv[i] = ub;
    }
}

// ------------------------------------------------------------------------------ one full cycle
struct CycleOut {
    int status = -1, iters = 0, n = 0, me = 0, mi = 0, iq = 0;
    Vec x, tau, dv, f, q_next, v_next, q_cmd;
    double momentum[3];
    std::vector<int> active; // CI row indices of the final active set (in activation order)
    // debug
    Mat H, CE, CI;
    Vec g, ce0, ci0;
    Dyn D;
};

void cycle(const iwbc_problem_t& P, const double* q, const double* v, const double* ref, unsigned contact_mask, CycleOut& O,
    bool keep_debug)
{
    const int nv = P.nv, na = P.na, nu = nv - na;
    Dyn& D = O.D;
    ORACLE_PHASE(PH_DYN);
    compute_all_terms(P, q, v, D);
    ORACLE_PHASE(PH_ASM);

    // active contacts and their force-column offsets
    int cidx[IWBC_MAXC], nk = 0;
    for (int c = 0; c < P.ncontacts; ++c) {
        if (contact_mask & (1u << c)) {
            cidx[c] = nk;
            nk += 12;
        }
        else
            cidx[c] = -1;
    }
    const int n = nv + nk;

    // ---- contacts: motion task, Jc = T^T J
    Mat Jc(nk, nv);
    std::vector<Mat> cA(P.ncontacts);
    double cb[IWBC_MAXC][6];
    for (int c = 0; c < P.ncontacts; ++c) {
        if (cidx[c] < 0)
            continue;
        const iwbc_contact_t& C = P.contacts[c];
        se3_task_rows(P, D, C.frame, C.kp, C.kd, ref + C.ref_off, cA[c], cb[c]);
        for (int r = 0; r < 12; ++r)
            for (int k = 0; k < nv; ++k) {
                double s = 0;
                for (int i = 0; i < 6; ++i)
                    s += C.T[i * 12 + r] * cA[c](i, k);
                Jc(cidx[c] + r, k) = s;
            }
    }

    // ---- HQP level 0 (constraints) in insertion order -> CE / CI as SolverHQuadProgFast stacks them
    int me = 0, mi = 0;
    for (int b = 0; b < P.nl0; ++b) {
        const iwbc_block_t& B = P.l0[b];
        if (B.kind == IWBC_B_BASEDYN)
            me += nu;
        else if (B.kind == IWBC_B_CONTACT_MOTION)
            me += cidx[B.index] >= 0 ? 6 : 0;
        else if (B.kind == IWBC_B_CONTACT_FORCE)
            mi += cidx[B.index] >= 0 ? 2 * 17 : 0;
        else if (B.kind == IWBC_B_TASK)
            mi += 2 * P.tasks[B.index].nrows;
    }
    Mat CE(me, n), CI(mi, n);
    Vec ce0(me, 0.0), ci0(mi, 0.0);
    int ie = 0, ii = 0;
    // removed-then-re-added contacts go to the END of level 0 in TSID; with a fixed mask we keep file order
    for (int b = 0; b < P.nl0; ++b) {
        const iwbc_block_t& B = P.l0[b];
        if (B.kind == IWBC_B_BASEDYN) {
            // [M_u | -J_u^T] x = -h_u  ->  CE = A, ce0 = -b = h_u
            for (int r = 0; r < nu; ++r) {
                for (int k = 0; k < nv; ++k)
                    CE(ie + r, k) = D.M(r, k);
                for (int k = 0; k < nk; ++k)
                    CE(ie + r, nv + k) = -Jc(k, r);
                ce0[ie + r] = D.nle[r];
            }
            ie += nu;
        }
        else if (B.kind == IWBC_B_CONTACT_MOTION) {
            int c = B.index;
            if (cidx[c] < 0)
                continue;
            for (int r = 0; r < 6; ++r) {
                for (int k = 0; k < nv; ++k)
                    CE(ie + r, k) = cA[c](r, k);
                ce0[ie + r] = -cb[c][r];
            }
            ie += 6;
        }
        else if (B.kind == IWBC_B_CONTACT_FORCE) {
            int c = B.index;
            if (cidx[c] < 0)
                continue;
            const iwbc_contact_t& C = P.contacts[c];
            for (int r = 0; r < 17; ++r) {
                for (int k = 0; k < 12; ++k) {
                    CI(ii + r, nv + cidx[c] + k) = C.B[r * 12 + k];
                    CI(ii + 17 + r, nv + cidx[c] + k) = -C.B[r * 12 + k];
                }
                ci0[ii + r] = -C.lb[r];
                ci0[ii + 17 + r] = C.ub[r];
            }
            ii += 34;
        }
        else if (B.kind == IWBC_B_TASK) {
            const iwbc_task_t& T = P.tasks[B.index];
            if (T.type == IWBC_T_BOUNDS) {
                Vec lb(na), ub(na);
                joint_acc_bounds(P, q, v, lb.data(), ub.data());
                for (int r = 0; r < na; ++r) {
                    CI(ii + r, nu + r) = 1.0;
                    CI(ii + na + r, nu + r) = -1.0;
                    ci0[ii + r] = -lb[r];
                    ci0[ii + na + r] = ub[r];
                }
                ii += 2 * na;
            }
            else if (T.type == IWBC_T_ACTBOUNDS) {
                // lb - h_a <= [M_a | -J_a^T] x <= ub - h_a
                for (int r = 0; r < na; ++r) {
                    for (int k = 0; k < nv; ++k) {
                        CI(ii + r, k) = D.M(nu + r, k);
                        CI(ii + na + r, k) = -D.M(nu + r, k);
                    }
                    for (int k = 0; k < nk; ++k) {
                        CI(ii + r, nv + k) = -Jc(k, nu + r);
                        CI(ii + na + r, nv + k) = Jc(k, nu + r);
                    }
                    ci0[ii + r] = -(-P.tau_max[r] - D.nle[nu + r]);
                    ci0[ii + na + r] = P.tau_max[r] - D.nle[nu + r];
                }
                ii += 2 * na;
            }
        }
    }

    // ---- HQP level 1 (costs): H = sum w A'A + reg I, g = -sum w A'b
    Mat H(n, n);
    Vec g(n, 0.0);
    auto add_cost = [&](double w, const Mat& A, const double* b, int rows) {
        ORACLE_PHASE(PH_QP); // SolverHQuadProgFast builds H and g from the level-1 rows
        for (int r = 0; r < rows; ++r) {
            for (int i = 0; i < n; ++i) {
                double ai = A(r, i);
                if (ai == 0.0)
                    continue;
                for (int j = 0; j < n; ++j)
                    H(i, j) += w * ai * A(r, j);
                g[i] -= w * ai * b[r];
            }
        }
        ORACLE_PHASE(PH_ASM);
    };
    for (int bl = 0; bl < P.nl1; ++bl) {
        const iwbc_block_t& B = P.l1[bl];
        if (B.kind == IWBC_B_CONTACT_REG) {
            int c = B.index;
            if (cidx[c] < 0)
                continue;
            const iwbc_contact_t& C = P.contacts[c];
            // Contact6d::updateForceRegularizationTask: matrix = diag(w) T, vector = diag(w) f_ref; the instance's reference row
            // carries f_ref (Contact6d::setForceReference) and which weight vector is installed (setRegularizationTaskWeightVector:
            // the stabilizer's zmp_w, src/controllers/humanoid_pos_tracker.cpp:220-223)
            const double* fref = ref + C.ref_off + IWBC_REF_CONTACT_FREF;
            int wv = (int)ref[C.ref_off + IWBC_REF_CONTACT_WREG];
            if (wv < 0 || wv >= P.nwreg)
                wv = 0;
            const real_t* wreg = P.wreg_variants[wv];
            Mat A(6, n);
            double b[6];
            for (int r = 0; r < 6; ++r) {
                for (int k = 0; k < 12; ++k)
                    A(r, nv + cidx[c] + k) = wreg[r] * C.T[r * 12 + k];
                b[r] = wreg[r] * fref[r];
            }
            add_cost(C.w_forcereg, A, b, 6);
            continue;
        }
        const iwbc_task_t& T = P.tasks[B.index];
        if (T.type == IWBC_T_SE3) {
            Mat A6;
            double b6[6];
            se3_task_rows(P, D, T.frame, T.kp, T.kd, ref + T.ref_off, A6, b6);
            Mat A(T.nrows, n);
            double b[6];
            int idx = 0;
            for (int i = 0; i < 6; ++i) {
                if (!T.mask[i])
                    continue;
                for (int k = 0; k < nv; ++k)
                    A(idx, k) = A6(i, k);
                b[idx++] = b6[i];
            }
            add_cost(T.weight, A, b, T.nrows);
        }
        else if (T.type == IWBC_T_COM) {
            const double* r = ref + T.ref_off;
            double com[3] = {D.com.x, D.com.y, D.com.z}, vc[3] = {D.vcom.x, D.vcom.y, D.vcom.z}, ac[3] = {D.acom.x, D.acom.y, D.acom.z};
            Mat A(T.nrows, n);
            double b[3];
            int idx = 0;
            for (int i = 0; i < 3; ++i) {
                if (!T.mask[i])
                    continue;
                double a_des = -T.kp[i] * (com[i] - r[i]) - T.kd[i] * (vc[i] - r[3 + i]) + r[6 + i];
                for (int k = 0; k < nv; ++k)
                    A(idx, k) = D.Jcom(i, k);
                b[idx++] = a_des - ac[i];
            }
            add_cost(T.weight, A, b, T.nrows);
        }
        else if (T.type == IWBC_T_POSTURE) {
            const double* r = ref + T.ref_off;
            Mat A(T.nrows, n);
            Vec b(T.nrows);
            int idx = 0;
            for (int i = 0; i < na; ++i) {
                if (!T.mask[i])
                    continue;
                double a_des = -T.kp[0] * (q[P.nq - na + i] - r[i]) - T.kd[0] * (v[nv - na + i] - r[na + i]) + r[2 * na + i];
                A(idx, nu + i) = 1.0;
                b[idx++] = a_des;
            }
            add_cost(T.weight, A, b.data(), T.nrows);
        }
        else if (T.type == IWBC_T_COP) {
            // tsid::tasks::TaskCopEquality added with addForceTask(task, w, 1) (tasks.cpp:156-178).  Upstream TSID is not in the reference
            // tree; restated from its published source: for every contact of the formulation's (active) contact list and every contact
            // point j, the three columns of the point's force coefficients get (p_world_j - cop_ref) * normal' (task-cop-equality.cpp:
            // compute()), the vector stays zero, and InverseDynamicsFormulationAccForce::computeProblemData writes that matrix into the
            // force columns of a level-1 constraint (m_taskContactForces loop: "by default the task is associated to all contact forces").
            const double* r = ref + T.ref_off; // cop reference (3), contact normal (3)
            Mat A(3, n);
            double b[3] = {0, 0, 0};
            for (int c = 0; c < P.ncontacts; ++c) {
                if (cidx[c] < 0)
                    continue;
                const iwbc_contact_t& C = P.contacts[c];
                FrameKin K;
                frame_kinematics(P, D, C.frame, K);
                const SE3& oMf = K.oMf; // Contact6d's motion-task frame: robot.framePosition(data, frame_id)
                for (int j = 0; j < 4; ++j) {
                    // contact point j in the contact frame (tasks.cpp:353-356), read back from T = [I ; skew(P_j)]
                    V3 Pl = {C.T[5 * 12 + 3 * j + 1], C.T[3 * 12 + 3 * j + 2], C.T[4 * 12 + 3 * j + 0]};
                    V3 pw = oMf.R * Pl + oMf.p;
                    const double pd[3] = {pw.x - r[0], pw.y - r[1], pw.z - r[2]};
                    for (int d = 0; d < 3; ++d)
                        for (int a = 0; a < 3; ++a)
                            A(d, nv + cidx[c] + 3 * j + a) = pd[d] * r[3 + a];
                }
            }
            add_cost(T.weight, A, b, 3);
        }
        else if (T.type == IWBC_T_TORQUE) {
            // tsid::tasks::TaskActuationEquality added with addActuationTask(task, w, 1) (tasks.cpp:227-271).  Upstream TSID is not in
            // the reference tree; restated from its published source: the task's own constraint is row i = w_i e_i', vector
            // w_i tau_ref_i over the masked actuated joints (task-actuation-equality.cpp: compute()), and
            // InverseDynamicsFormulationAccForce::computeProblemData turns an actuation constraint A_tau tau = b into rows over
            // [dv ; f] with tau = M_a dv + h_a - J_a^T f:  [A_tau M_a | -A_tau J_a^T] x = b - A_tau h_a.
            const double* r = ref + T.ref_off; // tau_ref (na), weight vector (na)
            Mat A(T.nrows, n);
            Vec b(T.nrows);
            int idx = 0;
            for (int i = 0; i < na; ++i) {
                if (!T.mask[i])
                    continue;
                const double s = r[na + i];
                for (int k = 0; k < nv; ++k)
                    A(idx, k) = s * D.M(nu + i, k);
                for (int k = 0; k < nk; ++k)
                    A(idx, nv + k) = -s * Jc(k, nu + i);
                b[idx++] = s * (r[i] - D.nle[nu + i]);
            }
            add_cost(T.weight, A, b.data(), T.nrows);
        }
        else if (T.type == IWBC_T_MOMENTUM) {
            const double* r = ref + T.ref_off;
            double L[6], drift[6] = {D.dhg.lin.x, D.dhg.lin.y, D.dhg.lin.z, D.dhg.ang.x, D.dhg.ang.y, D.dhg.ang.z};
            for (int i = 0; i < 6; ++i) {
                double s = 0;
                for (int k = 0; k < nv; ++k)
                    s += D.Ag(i, k) * v[k];
                L[i] = s;
            }
            Mat A(T.nrows, n);
            double b[6];
            int idx = 0;
            for (int i = 0; i < 6; ++i) {
                if (!T.mask[i])
                    continue;
                double dL_des = -T.kp[i] * (L[i] - r[6 + i]) + r[12 + i];
                for (int k = 0; k < nv; ++k)
                    A(idx, k) = D.Ag(i, k);
                b[idx++] = dL_des - drift[i];
            }
            add_cost(T.weight, A, b, T.nrows);
        }
        else if (T.type == IWBC_T_SELFCOL) {
            FrameKin K0;
            frame_kinematics(P, D, T.frame, K0, false);
            auto world_jac = [&](int frame, const SE3& oMf, Mat& Jw) {
                // 3 x nv linear part: WORLD (spatial, about the origin) or LOCAL_WORLD_ALIGNED
                Jw = Mat(3, nv);
                int pj = P.fparent[frame];
                if (pj < 0)
                    return;
                for (int j = 0; j < P.nj; ++j) {
                    if (!supports(P, j, pj))
                        continue;
                    for (int k = 0; k < joint_nv(P, j); ++k) {
                        int c = P.idx_v[j] + k;
                        V3 lin = {D.J(0, c), D.J(1, c), D.J(2, c)}, ang = {D.J(3, c), D.J(4, c), D.J(5, c)};
                        if (!P.sc_jacobian_world)
                            lin = lin + cross(ang, oMf.p);
                        Jw(0, c) = lin.x;
                        Jw(1, c) = lin.y;
                        Jw(2, c) = lin.z;
                    }
                }
            };
            Mat J1;
            world_jac(T.frame, K0.oMf, J1);
            Mat A(1, n);
            double Bv = 0;
            for (int a = 0; a < T.nav; ++a) {
                FrameKin Ka;
                frame_kinematics(P, D, T.av_frame[a], Ka, false);
                Mat Ja;
                world_jac(T.av_frame[a], Ka.oMf, Ja);
                V3 diff = K0.oMf.p - Ka.oMf.p;
                double sq = dot(diff, diff), nrm = std::sqrt(sq);
                double aa = T.av_r0[a] + T.radius;
                V3 drift = K0.a_classic.lin - Ka.a_classic.lin;
                double m = T.sc_m, k = T.sc_k;
                double x = k * (nrm - aa + T.sc_sp);
                double e_p = std::exp(-x);
                double Cv = 1. - std::pow(1 + e_p, -m);
                double gs = -1. / nrm * k * m * e_p * std::pow(e_p + 1, -m - 1);
                V3 grad = gs * diff;
                double h = 1. / sq * k * k * (-m - 1.) * m * std::exp(-2 * x) * std::pow(e_p + 1, -m - 2)
                    + 1. / sq * k * k * m * e_p * std::pow(e_p + 1, -m - 1)
                    + 1. / std::pow(nrm, 1.5) * k * m * e_p * std::pow(e_p + 1, -m - 1);
                // Hessian_C = h diff diff' + gs I
                V3 Jv = {0, 0, 0};
                for (int c = 0; c < nv; ++c) {
                    V3 col = {J1(0, c) - Ja(0, c), J1(1, c) - Ja(1, c), J1(2, c) - Ja(2, c)};
                    A(0, c) += dot(grad, col);
                    Jv = Jv + v[c] * col;
                }
                V3 HJv = (h * dot(diff, Jv)) * diff + gs * Jv;
                Bv += -(dot(HJv, Jv) + dot(grad, (-1.0 * drift) + T.kd[0] * Jv) + T.kp[0] * Cv);
            }
            add_cost(T.weight, A, &Bv, 1);
        }
    }
    for (int i = 0; i < n; ++i)
        H(i, i) += P.hess_reg;

    // ---- solve
    ORACLE_PHASE(PH_QP);
    QPWork w;
    Vec x;
    int st = solve_quadprog(H, g, CE, ce0, CI, ci0, x, w);
    // status map of SolverHQuadProgFast (Appendix A.6)
    O.status = st == QP_OPTIMAL ? IWBC_STATUS_OPTIMAL
        : st == QP_UNBOUNDED    ? IWBC_STATUS_INFEASIBLE
        : st == QP_MAX_ITER     ? IWBC_STATUS_MAX_ITER
        : st == QP_REDUNDANT_EQ ? IWBC_STATUS_ERROR
                                : IWBC_STATUS_UNKNOWN;
    O.iters = w.iter;
    O.n = n;
    O.me = me;
    O.mi = mi;
    O.iq = w.iq;
    O.x = x;
    O.active.clear();
    for (int i = me; i < w.iq; ++i)
        O.active.push_back(w.A[i]);

    // ---- decode (controller.cpp:250-260) and integrate (:254-256)
    ORACLE_PHASE(PH_DEC);
    O.dv.assign(x.begin(), x.begin() + nv);
    O.f.assign(12 * P.ncontacts, 0.0);
    for (int c = 0; c < P.ncontacts; ++c)
        if (cidx[c] >= 0)
            for (int k = 0; k < 12; ++k)
                O.f[12 * c + k] = x[nv + cidx[c] + k];
    O.tau.assign(na, 0.0);
    for (int r = 0; r < na; ++r) {
        double s = D.nle[nu + r];
        for (int k = 0; k < nv; ++k)
            s += D.M(nu + r, k) * x[k];
        for (int k = 0; k < nk; ++k)
            s -= Jc(k, nu + r) * x[nv + k];
        O.tau[r] = s;
    }
    O.v_next.resize(nv);
    for (int k = 0; k < nv; ++k)
        O.v_next[k] = v[k] + P.dt * x[k];
    O.q_next.assign(q, q + P.nq);
    for (int j = 0; j < P.nj; ++j) {
        if (P.jtype[j] == IWBC_J_FREEFLYER) {
            const double* qq = q + P.idx_q[j];
            const double* vv = O.v_next.data() + P.idx_v[j];
            SE3 M0 = {quat_to_R(qq[3], qq[4], qq[5], qq[6]), {qq[0], qq[1], qq[2]}};
            Motion nu6 = {{P.dt * vv[0], P.dt * vv[1], P.dt * vv[2]}, {P.dt * vv[3], P.dt * vv[4], P.dt * vv[5]}};
            SE3 E = exp6(nu6);
            SE3 M1 = M0 * E;
            // quaternion of the increment, composed on the right
            V3 w3 = nu6.ang;
            double t = norm(w3), s2, c2;
            if (t < 1.220703125e-4) {
                s2 = 0.5 - t * t / 48;
                c2 = 1 - t * t / 8;
            }
            else {
                s2 = std::sin(t / 2) / t;
                c2 = std::cos(t / 2);
            }
            double ex = s2 * w3.x, ey = s2 * w3.y, ez = s2 * w3.z, ew = c2;
            double x0 = qq[3], y0 = qq[4], z0 = qq[5], w0 = qq[6];
            double rx = w0 * ex + x0 * ew + y0 * ez - z0 * ey;
            double ry = w0 * ey - x0 * ez + y0 * ew + z0 * ex;
            double rz = w0 * ez + x0 * ey - y0 * ex + z0 * ew;
            double rw = w0 * ew - x0 * ex - y0 * ey - z0 * ez;
            if (rx * x0 + ry * y0 + rz * z0 + rw * w0 < 0) {
                rx = -rx;
                ry = -ry;
                rz = -rz;
                rw = -rw;
            }
            double nn = std::sqrt(rx * rx + ry * ry + rz * rz + rw * rw);
            double* o = O.q_next.data() + P.idx_q[j];
            o[0] = M1.p.x;
            o[1] = M1.p.y;
            o[2] = M1.p.z;
            o[3] = rx / nn;
            o[4] = ry / nn;
            o[5] = rz / nn;
            o[6] = rw / nn;
        }
        else
            O.q_next[P.idx_q[j]] = q[P.idx_q[j]] + P.dt * O.v_next[P.idx_v[j]];
    }
    // "DART" command layout (controller.cpp:263-282): free-flyer quaternion -> angle * axis
    O.q_cmd.assign(nv, 0.0);
    if (P.floating_base) {
        const double* o = O.q_next.data();
        double vx = o[3], vy = o[4], vz = o[5], ww = o[6];
        double nrm = std::sqrt(vx * vx + vy * vy + vz * vz);
        double ax = 1, ay = 0, az = 0, ang = 0;
        if (nrm != 0) {
            ang = 2 * std::atan2(nrm, std::fabs(ww));
            if (ww < 0)
                nrm = -nrm;
            ax = vx / nrm;
            ay = vy / nrm;
            az = vz / nrm;
        }
        O.q_cmd[0] = o[0];
        O.q_cmd[1] = o[1];
        O.q_cmd[2] = o[2];
        O.q_cmd[3] = ang * ax;
        O.q_cmd[4] = ang * ay;
        O.q_cmd[5] = ang * az;
        for (int k = 0; k < na; ++k)
            O.q_cmd[6 + k] = o[7 + k];
    }
    else
        for (int k = 0; k < nv; ++k)
            O.q_cmd[k] = O.q_next[k];
    // momentum_ = Ag.bottomRows(3) * dq (controller.cpp:245)
    for (int i = 0; i < 3; ++i) {
        double s = 0;
        for (int k = 0; k < nv; ++k)
            s += D.Ag(3 + i, k) * v[k];
        O.momentum[i] = s;
    }
    if (keep_debug) {
        O.H = H;
        O.g = g;
        O.CE = CE;
        O.ce0 = ce0;
        O.CI = CI;
        O.ci0 = ci0;
    }
}

} // namespace

// ======================================================================================= C API
extern "C" {

// ORACLE_COUNT_FLOPS build only: operation counts of the calling thread since the last reset, out[phase][op] with phases
// {other, dynamics, assembly, qp, decode} and ops {add/sub, mul, div, sqrt/trig/exp}; returns 0, or -1 in the plain build
int oracle_flop_counts(long* out20, int reset)
{
#ifdef ORACLE_COUNT_FLOPS
    for (int p = 0; p < flopcount::NPH; ++p)
        for (int o = 0; o < flopcount::NOP; ++o) {
            if (out20)
                out20[p * flopcount::NOP + o] = flopcount::counts[p][o];
            if (reset)
                flopcount::counts[p][o] = 0;
        }
    flopcount::phase = 0;
    return 0;
#else
    (void)out20;
    (void)reset;
    return -1;
#endif
}

// Model collision check, restated from src/safety/collision_check.cpp:27-87: sphere centres = oMi[joint] * p; every ordered pair of
// DIFFERENT members, members in index order (the caller sorts them by name as the reference's std::map does), first hit returns.
// spheres: p_radius4[s] = {x, y, z, radius}, joint_member2[s] = {joint (-1 universe), member}.  out5 = {hit, member a, i, member b, j}.
int oracle_collision(const iwbc_problem_t* P, const double* q, int nspheres, const double* p_radius4, const int* joint_member2, int nmembers, int* out5)
{
    Dyn D;
    std::vector<double> v0(P->nv, 0.0);
    compute_all_terms(*P, q, v0.data(), D);
    std::vector<std::vector<std::pair<V3, double>>> members(nmembers);
    for (int s = 0; s < nspheres; ++s) {
        V3 rel = {p_radius4[4 * s], p_radius4[4 * s + 1], p_radius4[4 * s + 2]};
        const int j = joint_member2[2 * s];
        V3 c = j >= 0 ? D.oMi[j].R * rel + D.oMi[j].p : rel;
        members[joint_member2[2 * s + 1]].push_back(std::make_pair(c, p_radius4[4 * s + 3]));
    }
    out5[0] = 0, out5[1] = out5[2] = out5[3] = out5[4] = -1;
    for (int a = 0; a < nmembers; ++a)
        for (int b = 0; b < nmembers; ++b) {
            if (a == b)
                continue;
            for (size_t i = 0; i < members[a].size(); ++i)
                for (size_t j = 0; j < members[b].size(); ++j)
                    if (norm(members[b][j].first - members[a][i].first) < members[b][j].second + members[a][i].second) {
                        out5[0] = 1, out5[1] = a, out5[2] = (int)i, out5[3] = b, out5[4] = (int)j;
                        return 0;
                    }
        }
    return 0;
}

// counters of the Goldfarb-Idnani loop since the last reset: {l1 passes, step directions, drops} (development aid)
void oracle_qp_stats(long* out3, int reset)
{
    for (int k = 0; k < 3; ++k) {
        if (out3)
            out3[k] = g_stat[k + 1].load();
        if (reset)
            g_stat[k + 1] = 0;
    }
}


#define ORACLE_ACTIVE_WORDS 12

// Batch of B independent control updates, `nthreads` worker threads, instances statically striped
// over threads (the reference's own test harness model, tests/utest.hpp:79-92).
// Layout (instance-major): q[B][nq] v[B][nv] refs[B][nref] (or refs_stride = 0: one shared vector),
// contact_mask[B] or NULL (= all default-active contacts).
// Outputs (any may be NULL): tau[B][na] dv[B][nv] f[B][12*ncontacts] q_next[B][nq] v_next[B][nv]
// status[B] iters[B] active[B][ORACLE_ACTIVE_WORDS] (bitmask over CI rows) q_cmd[B][nv] momentum[B][3].
int oracle_cycle_batch(const iwbc_problem_t* P, int B, const double* q, const double* v, const double* refs, long refs_stride,
    const unsigned* contact_mask, double* tau, double* dv, double* f, double* q_next, double* v_next, int* status, int* iters,
    unsigned* active, double* q_cmd, double* momentum, int nthreads)
{
    if (!P || P->magic != IWBC_PROBLEM_MAGIC)
        return -1;
    unsigned dflt = 0;
    for (int c = 0; c < P->ncontacts; ++c)
        if (P->contacts[c].active_default)
            dflt |= 1u << c;
    auto work = [&](int t) {
        CycleOut O;
        for (int i = t; i < B; i += nthreads) {
            unsigned mask = contact_mask ? contact_mask[i] : dflt;
            cycle(*P, q + (size_t)i * P->nq, v + (size_t)i * P->nv, refs + (size_t)i * refs_stride, mask, O, false);
            if (tau)
                std::copy(O.tau.begin(), O.tau.end(), tau + (size_t)i * P->na);
            if (dv)
                std::copy(O.dv.begin(), O.dv.end(), dv + (size_t)i * P->nv);
            if (f)
                std::copy(O.f.begin(), O.f.end(), f + (size_t)i * 12 * P->ncontacts);
            if (q_next)
                std::copy(O.q_next.begin(), O.q_next.end(), q_next + (size_t)i * P->nq);
            if (v_next)
                std::copy(O.v_next.begin(), O.v_next.end(), v_next + (size_t)i * P->nv);
            if (q_cmd)
                std::copy(O.q_cmd.begin(), O.q_cmd.end(), q_cmd + (size_t)i * P->nv);
            if (momentum)
                std::copy(O.momentum, O.momentum + 3, momentum + (size_t)i * 3);
            if (status)
                status[i] = O.status;
            if (iters)
                iters[i] = O.iters;
            if (active) {
                unsigned* a = active + (size_t)i * ORACLE_ACTIVE_WORDS;
                std::memset(a, 0, sizeof(unsigned) * ORACLE_ACTIVE_WORDS);
                for (int r : O.active)
                    if (r >= 0 && r < 32 * ORACLE_ACTIVE_WORDS)
                        a[r >> 5] |= 1u << (r & 31);
            }
        }
    };
    if (nthreads <= 1) {
        nthreads = 1;
        work(0);
    }
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back(work, t);
        for (auto& t : th)
            t.join();
    }
    return 0;
}

// Open-loop roll-out of ONE controller for `ticks` ticks (Controller::update with closed_loop = false):
// state (q, v) is advanced in place; refs_t[ticks][nref] (stride 0 = constant refs).  q_cmd_out[ticks][nv].
int oracle_rollout(const iwbc_problem_t* P, int ticks, double* q, double* v, const double* refs_t, long refs_stride,
    unsigned contact_mask, double* q_cmd_out, double* tau_out, int* status_out)
{
    CycleOut O;
    for (int t = 0; t < ticks; ++t) {
        cycle(*P, q, v, refs_t + (size_t)t * refs_stride, contact_mask, O, false);
        if (status_out)
            status_out[t] = O.status;
        if (O.status != IWBC_STATUS_OPTIMAL)
            return t + 1; // the reference throws here (controller.cpp:284-308)
        std::copy(O.q_next.begin(), O.q_next.end(), q);
        std::copy(O.v_next.begin(), O.v_next.end(), v);
        if (q_cmd_out)
            std::copy(O.q_cmd.begin(), O.q_cmd.end(), q_cmd_out + (size_t)t * P->nv);
        if (tau_out)
            std::copy(O.tau.begin(), O.tau.end(), tau_out + (size_t)t * P->na);
    }
    return 0;
}

// Single-instance debug dump for the stage-by-stage tests.  All matrices row-major; any pointer may be NULL.
// sizes[0..2] = n, me, mi.
int oracle_cycle_debug(const iwbc_problem_t* P, const double* q, const double* v, const double* ref, unsigned contact_mask,
    double* M, double* nle, double* Jworld, double* Jcom, double* Ag, double* com9 /* com, vcom, acom */,
    double* hg12 /* hg(6), dhg(6) */, double* H, double* g, double* CE, double* ce0, double* CI, double* ci0, double* x, int* sizes,
    int* status, double* oMi /* nj x 12 */)
{
    CycleOut O;
    cycle(*P, q, v, ref, contact_mask, O, true);
    auto cp = [](const Vec& a, double* o) {
        if (o)
            std::copy(a.begin(), a.end(), o);
    };
    cp(O.D.M.a, M);
    cp(O.D.nle, nle);
    cp(O.D.J.a, Jworld);
    cp(O.D.Jcom.a, Jcom);
    cp(O.D.Ag.a, Ag);
    if (com9) {
        double t[9] = {O.D.com.x, O.D.com.y, O.D.com.z, O.D.vcom.x, O.D.vcom.y, O.D.vcom.z, O.D.acom.x, O.D.acom.y, O.D.acom.z};
        std::copy(t, t + 9, com9);
    }
    if (hg12) {
        double t[12] = {O.D.hg.lin.x, O.D.hg.lin.y, O.D.hg.lin.z, O.D.hg.ang.x, O.D.hg.ang.y, O.D.hg.ang.z, O.D.dhg.lin.x,
            O.D.dhg.lin.y, O.D.dhg.lin.z, O.D.dhg.ang.x, O.D.dhg.ang.y, O.D.dhg.ang.z};
        std::copy(t, t + 12, hg12);
    }
    cp(O.H.a, H);
    cp(O.g, g);
    cp(O.CE.a, CE);
    cp(O.ce0, ce0);
    cp(O.CI.a, CI);
    cp(O.ci0, ci0);
    cp(O.x, x);
    if (sizes) {
        sizes[0] = O.n;
        sizes[1] = O.me;
        sizes[2] = O.mi;
    }
    if (status)
        *status = O.status;
    if (oMi)
        for (int j = 0; j < P->nj; ++j) {
            for (int k = 0; k < 9; ++k)
                oMi[12 * j + k] = O.D.oMi[j].R.m[k];
            oMi[12 * j + 9] = O.D.oMi[j].p.x;
            oMi[12 * j + 10] = O.D.oMi[j].p.y;
            oMi[12 * j + 11] = O.D.oMi[j].p.z;
        }
    return 0;
}

// Generic dense QP (row-major H n x n, CE me x n, CI mi x n), eiquadprog-fast conventions.
// Returns the eiquadprog status (0 optimal, 1 infeasible(unbounded dual), 3 max-iter, 4 redundant equalities).
int oracle_qp_solve(int n, int me, int mi, const double* H, const double* g, const double* CE, const double* ce0,
    const double* CI, const double* ci0, double* x, int* active /* mi entries, -1 terminated */, int* iters, double* lambda /* me+mi */)
{
    Mat Hm(n, n), CEm(me, n), CIm(mi, n);
    std::copy(H, H + (size_t)n * n, Hm.a.begin());
    if (me)
        std::copy(CE, CE + (size_t)me * n, CEm.a.begin());
    if (mi)
        std::copy(CI, CI + (size_t)mi * n, CIm.a.begin());
    Vec gv(g, g + n), ce(ce0, ce0 + me), ci(ci0, ci0 + mi), xv;
    QPWork w;
    int st = solve_quadprog(Hm, gv, CEm, ce, CIm, ci, xv, w);
    std::copy(xv.begin(), xv.end(), x);
    if (iters)
        *iters = w.iter;
    if (active) {
        int k = 0;
        for (int i = me; i < w.iq && k < mi; ++i)
            active[k++] = w.A[i];
        if (k < mi)
            active[k] = -1;
    }
    if (lambda) {
        for (int i = 0; i < me + mi; ++i)
            lambda[i] = 0;
        for (int i = 0; i < w.iq; ++i) {
            int a = w.A[i];
            if (a < 0)
                lambda[-a - 1] = w.u[i];
            else
                lambda[me + a] = w.u[i];
        }
    }
    return st == QP_OPTIMAL ? 0 : st == QP_UNBOUNDED ? 1 : st == QP_MAX_ITER ? 3 : 4;
}

// RNEA with arbitrary joint acceleration: independent check of CRBA / nle (M a + nle == rnea(q, v, a)).
int oracle_rnea(const iwbc_problem_t* P, const double* q, const double* vq, const double* aq, double* tau)
{
    const int nj = P->nj;
    std::vector<SE3> liMi(nj);
    std::vector<Motion> v(nj), a(nj);
    std::vector<Force> f(nj);
    const Motion zero = {{0, 0, 0}, {0, 0, 0}};
    const Motion a0 = {{-P->gravity[0], -P->gravity[1], -P->gravity[2]}, {0, 0, 0}};
    for (int j = 0; j < nj; ++j) {
        SE3 jM = {identity3(), {0, 0, 0}};
        V3 ax = {P->axis[j][0], P->axis[j][1], P->axis[j][2]};
        Motion vJ = zero, aJ = zero;
        if (P->jtype[j] == IWBC_J_FREEFLYER) {
            const double* qq = q + P->idx_q[j];
            jM.R = quat_to_R(qq[3], qq[4], qq[5], qq[6]);
            jM.p = {qq[0], qq[1], qq[2]};
            const double* vv = vq + P->idx_v[j];
            const double* aa = aq + P->idx_v[j];
            vJ = {{vv[0], vv[1], vv[2]}, {vv[3], vv[4], vv[5]}};
            aJ = {{aa[0], aa[1], aa[2]}, {aa[3], aa[4], aa[5]}};
        }
        else if (P->jtype[j] == IWBC_J_REVOLUTE) {
            jM.R = axis_angle(ax, q[P->idx_q[j]]);
            vJ.ang = vq[P->idx_v[j]] * ax;
            aJ.ang = aq[P->idx_v[j]] * ax;
        }
        else {
            jM.p = q[P->idx_q[j]] * ax;
            vJ.lin = vq[P->idx_v[j]] * ax;
            aJ.lin = aq[P->idx_v[j]] * ax;
        }
        liMi[j] = load_se3(P->jplace[j]) * jM;
        int p = P->parent[j];
        v[j] = (p >= 0 ? actInv(liMi[j], v[p]) : zero) + vJ;
        a[j] = actInv(liMi[j], p >= 0 ? a[p] : a0) + aJ + mcross(v[j], vJ);
        Inertia Y = joint_inertia(*P, j);
        f[j] = Y * a[j] + mcrossf(v[j], Y * v[j]);
    }
    for (int j = nj - 1; j >= 0; --j) {
        for (int k = 0; k < joint_nv(*P, j); ++k) {
            Motion S = subspace(*P, j, k);
            tau[P->idx_v[j] + k] = dot(S.lin, f[j].lin) + dot(S.ang, f[j].ang);
        }
        if (P->parent[j] >= 0)
            f[P->parent[j]] = f[P->parent[j]] + act(liMi[j], f[j]);
    }
    return 0;
}

// frame placement / local velocity / classic acceleration / local Jacobian of task frame `frame`
int oracle_frame(const iwbc_problem_t* P, const double* q, const double* v, int frame, double* oMf12, double* vel6, double* acc6, double* Jlocal)
{
    Dyn D;
    compute_all_terms(*P, q, v, D);
    FrameKin K;
    frame_kinematics(*P, D, frame, K);
    if (oMf12) {
        for (int k = 0; k < 9; ++k)
            oMf12[k] = K.oMf.R.m[k];
        oMf12[9] = K.oMf.p.x;
        oMf12[10] = K.oMf.p.y;
        oMf12[11] = K.oMf.p.z;
    }
    if (vel6) {
        double t[6] = {K.v.lin.x, K.v.lin.y, K.v.lin.z, K.v.ang.x, K.v.ang.y, K.v.ang.z};
        std::copy(t, t + 6, vel6);
    }
    if (acc6) {
        double t[6] = {K.a_classic.lin.x, K.a_classic.lin.y, K.a_classic.lin.z, K.a_classic.ang.x, K.a_classic.ang.y, K.a_classic.ang.z};
        std::copy(t, t + 6, acc6);
    }
    if (Jlocal)
        std::copy(K.Jlocal.a.begin(), K.Jlocal.a.end(), Jlocal);
    return 0;
}

// What the stabilizer reads from tsid's data (src/controllers/humanoid_pos_tracker.cpp:170-176,205; talos_pos_tracker.cpp:142):
// out21 = com(3) vcom(3) acom(3) [data.com[0], vcom[0], acom[0] after computeAllTerms + centerOfMass(q, v, 0)], the placements'
// translations of the two contact joints (contact_lfoot, contact_rfoot), the angular velocity of imu_link in its own frame.
int oracle_stab_model(const iwbc_problem_t* P, const double* q, const double* v, double* out21)
{
    Dyn D;
    compute_all_terms(*P, q, v, D);
    double t[21] = {D.com.x, D.com.y, D.com.z, D.vcom.x, D.vcom.y, D.vcom.z, D.acom.x, D.acom.y, D.acom.z};
    for (int k = 9; k < 21; ++k)
        t[k] = 0;
    const int cs[2] = {P->stab_contact_lf, P->stab_contact_rf};
    for (int s = 0; s < 2; ++s)
        if (cs[s] >= 0) {
            FrameKin K;
            frame_kinematics(*P, D, P->contacts[cs[s]].frame, K);
            t[9 + 3 * s] = K.oMf.p.x, t[10 + 3 * s] = K.oMf.p.y, t[11 + 3 * s] = K.oMf.p.z;
        }
    if (P->stab_imu_frame >= 0) {
        FrameKin K;
        frame_kinematics(*P, D, P->stab_imu_frame, K);
        t[15] = K.v.ang.x, t[16] = K.v.ang.y, t[17] = K.v.ang.z;
    }
    std::copy(t, t + 21, out21);
    return 0;
}

int oracle_problem_size(void) { return (int)sizeof(iwbc_problem_t); }

} // extern "C"
