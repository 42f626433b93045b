// stabilizer.cu -- the reference's sensor-driven pre-step as a batched kernel pair.
//
// HumanoidPosTracker::update (src/controllers/humanoid_pos_tracker.cpp:134-302) wraps every solve: it estimates the centre of
// pressure from the feet's force / torque sensors (src/estimators/cop.cpp:10-96) through moving-average filters
// (include/inria_wbc/estimators/filtering.hpp:17-146), rewrites task references with four admittance laws
// (src/stabilizers/stabilizer.cpp: com_admittance :31-59, zmp_distributor_admittance :277-311 / :340-429, ankle_admittance :168-224,
// foot_force_difference_admittance :226-275), solves, and puts the references back (:291-301).  TalosPosTracker::update adds the
// IMU momentum admittance in front (src/controllers/talos_pos_tracker.cpp:135-150, stabilizer.cpp:107-140).
//
// Batched form: every robot of the batch owns its filter windows and the few values that survive a tick (normal forces of the last
// solve, the contacts' force references, the installed regularisation weight vector, the state tsid's data was computed on);
//   k_stab_pre   one thread per robot: kinematics of the data state (CoM sample, ankle placements, IMU-frame velocity), CoP
//                estimate, filters, the admittance laws -> a MODIFIED COPY of the robot's reference row (the caller's row is
//                never touched, which is the reference's "put back"), and the robot's contact-mask bucket;
//   [the control cycle on the modified rows, one shape-uniform launch set per contact mask]
//   k_stab_post  one thread per robot: normal forces of this solve and the data state for the next tick.
// The ZMP distributor gives every contact its own force reference and installs zmp_w as regularisation weights: the reference row
// carries both (iwbc_problem.h: contact block), the kernels take the matching host-factorised force block.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "batch_state.hpp"

using namespace iwbc;

#define IWBC_STAB_NF 8 // filters per robot: cop, lcop, rcop (2 values each); lf_force, rf_force, lf_torque, rf_torque, imu (3 each)

struct iwbc_stabilizer {
    iwbc_batch* batch = nullptr;
    int device = 0, cap = 0, W = 0; // robots, longest filter window of the three parameter sets
    // per-robot state, structure of arrays (index ... * cap + robot)
    double* d_buf = nullptr;      // [IWBC_STAB_NF][3][W][cap] filter windows
    int* d_cnt = nullptr;         // [IWBC_STAB_NF][cap] samples seen (Filter::_cnt)
    int* d_wsize = nullptr;       // [cap] window size in use (all of a robot's filters share it)
    int* d_btype = nullptr;       // [cap] behaviour type the windows were last sized for
    double* d_ffilt = nullptr;    // [6][cap] _lf_force_filtered, _rf_force_filtered
    double* d_fn_prev = nullptr;  // [2][cap] normal forces of the last OPTIMAL solve (left, right)
    int* d_have_prev = nullptr;   // [cap] contact mask of that solve (activated_contacts_forces_'s keys), 0 = none yet
    int* d_wreg = nullptr;        // [cap] installed regularisation weight vector
    double* d_base_fref = nullptr; // [12][cap] the contacts' own force references
    double* d_qdata = nullptr;    // [cap][nq] the state tsid's data was last computed on
    double* d_vdata = nullptr;    // [cap][nv]
    double* d_refs = nullptr;     // [cap][nref] modified rows
    int* d_blist = nullptr;       // [3][cap] contact-mask buckets
    int* d_bcount = nullptr;      // [4]
    unsigned masks[3] = {0, 0, 0};
};

namespace {

struct StabIO {
    int B, cap, W, nref;
    long refs_stride;
    const double* refs;
    const unsigned* mask;   // per robot or NULL
    unsigned mask_all;      // used when mask == NULL
    const int* btype;       // per robot or NULL
    int btype_all;
    const double* sensors;  // [B][15]
    double* buf;
    int* cnt;
    int* wsize;
    int* btype_state;
    double* ffilt;
    double* fn_prev;
    int* have_prev;
    int* wreg;
    double* base_fref;
    const double* qdata;
    const double* vdata;
    double* out;            // [B][nref]
    int* blist;
    int* bcount;
    unsigned m0, m1, m2;
    int* status;            // rows with an unusable contact mask get IWBC_STATUS_ERROR (may be NULL)
};

struct V3d {
    double x, y, z;
};
__device__ __forceinline__ V3d mk3(double x, double y, double z) { return V3d{x, y, z}; }
__device__ __forceinline__ V3d add3(V3d a, V3d b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3d sub3(V3d a, V3d b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3d scl3(double s, V3d a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3d crs3(V3d a, V3d b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3d rot3(const double* R, V3d v)
{
    return mk3(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z);
}
__device__ __forceinline__ V3d rott3(const double* R, V3d v)
{
    return mk3(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z);
}
__device__ __forceinline__ void mul33(const double* A, const double* B, double* C)
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

// What the stabilizer reads from tsid's data: data.com[0] / vcom[0] / acom[0] (pinocchio computeAllTerms + centerOfMass(q, v, 0):
// drift acceleration, no gravity), the translation of the two contact joints' placements, the angular velocity of imu_link in its
// own frame.  World-frame point kinematics along the tree; joints are numbered depth first, so the chain of ancestors of the
// current joint is always alive in a stack indexed by depth.
struct Kin {
    double R[9];
    V3d p, w, v, al, a; // placement; angular / linear velocity of the origin; angular / linear (classical) drift acceleration
};
#define IWBC_STAB_MAXD 24
__device__ void stab_model(const iwbc_problem_t* __restrict__ P, const double* __restrict__ q, const double* __restrict__ v, double* com9, V3d& lf_p,
    V3d& rf_p, V3d& imu_w)
{
    Kin st[IWBC_STAB_MAXD];
    V3d mc = mk3(0, 0, 0), mv = mk3(0, 0, 0), ma = mk3(0, 0, 0);
    double mass = 0;
    const int jl = P->stab_contact_lf >= 0 ? P->fparent[P->contacts[P->stab_contact_lf].frame] : -2;
    const int jr = P->stab_contact_rf >= 0 ? P->fparent[P->contacts[P->stab_contact_rf].frame] : -2;
    const int jimu = P->stab_imu_frame >= 0 ? P->fparent[P->stab_imu_frame] : -2;
    lf_p = rf_p = imu_w = mk3(0, 0, 0);
    for (int j = 0; j < P->nj; ++j) {
        const int d = P->depth[j], ty = P->jtype[j];
        Kin& k = st[d];
        const double* jp = P->jplace[j];
        const double* qq = q + P->idx_q[j];
        const double* vv = v + P->idx_v[j];
        // parent (world) quantities
        double Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        V3d pp = mk3(0, 0, 0), wp = pp, vp = pp, alp = pp, ap = pp;
        if (P->parent[j] >= 0) {
            const Kin& par = st[d - 1];
#pragma unroll
            for (int e = 0; e < 9; ++e)
                Rp[e] = par.R[e];
            pp = par.p, wp = par.w, vp = par.v, alp = par.al, ap = par.a;
        }
        double Rj[9]; // joint rotation, and the joint's own motion expressed in the frame of jplace
        V3d tj = mk3(0, 0, 0);
        if (ty == IWBC_J_FREEFLYER) {
            const double x = qq[3], y = qq[4], z = qq[5], w = qq[6];
            Rj[0] = 1 - 2 * (y * y + z * z), Rj[1] = 2 * (x * y - z * w), Rj[2] = 2 * (x * z + y * w);
            Rj[3] = 2 * (x * y + z * w), Rj[4] = 1 - 2 * (x * x + z * z), Rj[5] = 2 * (y * z - x * w);
            Rj[6] = 2 * (x * z - y * w), Rj[7] = 2 * (y * z + x * w), Rj[8] = 1 - 2 * (x * x + y * y);
            tj = mk3(qq[0], qq[1], qq[2]);
        }
        else if (ty == IWBC_J_REVOLUTE) {
            const double ax = P->axis[j][0], ay = P->axis[j][1], az = P->axis[j][2];
            double s, c;
            sincos(qq[0], &s, &c);
            const double oc = 1 - c;
            Rj[0] = c + ax * ax * oc, Rj[1] = ax * ay * oc - az * s, Rj[2] = ax * az * oc + ay * s;
            Rj[3] = ay * ax * oc + az * s, Rj[4] = c + ay * ay * oc, Rj[5] = ay * az * oc - ax * s;
            Rj[6] = az * ax * oc - ay * s, Rj[7] = az * ay * oc + ax * s, Rj[8] = c + az * az * oc;
        }
        else {
            Rj[0] = Rj[4] = Rj[8] = 1, Rj[1] = Rj[2] = Rj[3] = Rj[5] = Rj[6] = Rj[7] = 0;
            tj = scl3(qq[0], mk3(P->axis[j][0], P->axis[j][1], P->axis[j][2]));
        }
        double Rpl[9]; // world rotation of the jplace frame
        mul33(Rp, jp, Rpl);
        mul33(Rpl, Rj, k.R);
        const V3d r = add3(rot3(Rp, mk3(jp[9], jp[10], jp[11])), rot3(Rpl, tj)); // parent origin -> this joint's origin, world axes
        k.p = add3(pp, r);
        // the parent body carries the point `r`: velocity and classical acceleration of that point
        V3d vpt = add3(vp, crs3(wp, r));
        V3d apt = add3(ap, add3(crs3(alp, r), crs3(wp, crs3(wp, r))));
        if (ty == IWBC_J_FREEFLYER) {
            // v = [linear; angular] in the body frame; zero joint acceleration: the body's spatial acceleration is its parent's, the
            // classical acceleration of the origin picks up w x v
            const V3d vl = rot3(k.R, mk3(vv[0], vv[1], vv[2])), wl = rot3(k.R, mk3(vv[3], vv[4], vv[5]));
            k.w = add3(wp, wl);
            k.v = add3(vpt, vl);
            k.al = add3(alp, crs3(wp, wl));
            k.a = add3(apt, add3(scl3(2.0, crs3(wp, vl)), crs3(wl, vl)));
        }
        else if (ty == IWBC_J_REVOLUTE) {
            const V3d wj = scl3(vv[0], rot3(k.R, mk3(P->axis[j][0], P->axis[j][1], P->axis[j][2])));
            k.w = add3(wp, wj);
            k.v = vpt;
            k.al = add3(alp, crs3(wp, wj));
            k.a = apt;
        }
        else {
            const V3d vj = scl3(vv[0], rot3(Rpl, mk3(P->axis[j][0], P->axis[j][1], P->axis[j][2])));
            k.w = wp;
            k.v = add3(vpt, vj);
            k.al = alp;
            k.a = add3(apt, scl3(2.0, crs3(wp, vj)));
        }
        // the body's centre of mass
        const double m = P->mass[j];
        const V3d rc = rot3(k.R, mk3(P->com[j][0], P->com[j][1], P->com[j][2]));
        mc = add3(mc, scl3(m, add3(k.p, rc)));
        mv = add3(mv, scl3(m, add3(k.v, crs3(k.w, rc))));
        ma = add3(ma, scl3(m, add3(k.a, add3(crs3(k.al, rc), crs3(k.w, crs3(k.w, rc))))));
        mass += m;
        if (j == jl)
            lf_p = k.p;
        if (j == jr)
            rf_p = k.p;
        if (j == jimu) {
            double Rf[9];
            mul33(k.R, P->fplace[P->stab_imu_frame], Rf);
            imu_w = rott3(Rf, k.w);
        }
    }
    const double im = 1.0 / mass;
    com9[0] = mc.x * im, com9[1] = mc.y * im, com9[2] = mc.z * im;
    com9[3] = mv.x * im, com9[4] = mv.y * im, com9[5] = mv.z * im;
    com9[6] = ma.x * im, com9[7] = ma.y * im, com9[8] = ma.z * im;
}

// ---- Filter::filter + MovingAverageFilter::_filter_impl (filtering.hpp:41-61, 139-142) on the robot's window f
__device__ void filter_push(const StabIO& s, int i, int f, int nvar, const double* sample, double* filtered)
{
    const int w = s.wsize[i];
    int cnt = s.cnt[(size_t)f * s.cap + i];
    for (int var = 0; var < nvar; ++var) {
        double* col = s.buf + ((size_t)(f * 3 + var) * s.W) * s.cap + i; // element c at col[c * cap]
        double acc = 0;
        if (cnt < w) {
            col[(size_t)cnt * s.cap] = sample[var];
            for (int c = 0; c <= cnt; ++c)
                acc += col[(size_t)c * s.cap];
            filtered[var] = acc / (cnt + 1);
        }
        else {
            for (int c = 0; c + 1 < w; ++c)
                col[(size_t)c * s.cap] = col[(size_t)(c + 1) * s.cap];
            col[(size_t)(w - 1) * s.cap] = sample[var];
            for (int c = 0; c < w; ++c)
                acc += col[(size_t)c * s.cap];
            filtered[var] = acc / w;
        }
    }
    if (cnt < w)
        s.cnt[(size_t)f * s.cap + i] = cnt + 1;
}
__device__ void filter_reset(const StabIO& s, int i, int f)
{ // Filter::reset (filtering.hpp:88-98)
    for (int var = 0; var < 3; ++var)
        for (int c = 0; c < s.W; ++c)
            s.buf[((size_t)(f * 3 + var) * s.W + c) * s.cap + i] = 0.0;
    s.cnt[(size_t)f * s.cap + i] = 0;
}
// Filter::set_window_size (filtering.hpp:63-86): grow = keep the old columns in front, shrink = keep the last ones; _cnt untouched
__device__ void filter_resize(const StabIO& s, int i, int f, int w_old, int w_new)
{
    for (int var = 0; var < 3; ++var) {
        double* col = s.buf + ((size_t)(f * 3 + var) * s.W) * s.cap + i;
        if (w_new > w_old) {
            for (int c = w_old; c < w_new; ++c)
                col[(size_t)c * s.cap] = 0.0;
        }
        else {
            for (int c = 0; c < w_new; ++c)
                col[(size_t)c * s.cap] = col[(size_t)(c + w_old - w_new) * s.cap];
        }
    }
}

// Eigen::Matrix3d::eulerAngles(0, 1, 2) (Eigen 3.3 Geometry/EulerAngles.h: i, j, k = 0, 1, 2, odd = false); R row-major
__device__ void euler012(const double* R, double* e)
{
    double r0 = atan2(R[5], R[8]);
    const double c2 = sqrt(R[0] * R[0] + R[1] * R[1]);
    double r1;
    if (r0 > 0.0) {
        r0 -= 3.14159265358979323846;
        r1 = atan2(-R[2], -c2);
    }
    else
        r1 = atan2(-R[2], c2);
    double s1, c1;
    sincos(r0, &s1, &c1);
    const double r2 = atan2(s1 * R[6] - c1 * R[3], c1 * R[4] - s1 * R[7]);
    e[0] = -r0, e[1] = -r1, e[2] = -r2;
}
// (AngleAxis(e0, X) * AngleAxis(e1, Y) * AngleAxis(e2, Z)).toRotationMatrix(), written into a sample's column-major rotation block
__device__ void rot_xyz_to_sample(const double* e, double* pos12)
{
    double s0, c0, s1, c1, s2, c2;
    sincos(e[0], &s0, &c0);
    sincos(e[1], &s1, &c1);
    sincos(e[2], &s2, &c2);
    const double Rx[9] = {1, 0, 0, 0, c0, -s0, 0, s0, c0}, Ry[9] = {c1, 0, s1, 0, 1, 0, -s1, 0, c1}, Rz[9] = {c2, -s2, 0, s2, c2, 0, 0, 0, 1};
    double T[9], R[9];
    mul33(Rx, Ry, T);
    mul33(T, Rz, R);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            pos12[3 + 3 * c + r] = R[3 * r + c];
}
__device__ void sample_rotation(const double* pos12, double* R)
{ // vectorToSE3: column-major 3 x 3 after the translation
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            R[3 * r + c] = pos12[3 + 3 * c + r];
}

// stabilizer.cpp:340-429
__device__ void zmp_distributor(double mass, double zx, double zy, const double* lf_ref_p, const double* rf_ref_p, bool lf_active, bool rf_active,
    double* left, double* right, bool& ok)
{
    double alpha = 0, PRx = 0, PRy = 0, PLx = 0, PLy = 0;
    ok = true;
    if (lf_active && rf_active) {
        const double fx = lf_ref_p[0], fy = lf_ref_p[1], sx = rf_ref_p[0], sy = rf_ref_p[1]; // fline.first = left, .second = right, z = 0
        PRx = sx, PRy = sy, PLx = fx, PLy = fy;
        const double dx = sx - fx, dy = sy - fy, len = sqrt(dx * dx + dy * dy + 0.0);
        const double ux = dx / len, uy = dy / len; // closest_point_on_line (:431-440)
        const double dd = (zx - sx) * ux + (zy - sy) * uy + 0.0;
        const double px = sx + ux * dd, py = sy + uy * dd;
        const double n1 = sqrt((px - fx) * (px - fx) + (py - fy) * (py - fy) + 0.0), n2 = sqrt((sx - px) * (sx - px) + (sy - py) * (sy - py) + 0.0);
        if (fabs(len - n1 - n2) < 1e-10)
            alpha = n1 / len;
        else if (n1 < n2)
            alpha = 0;
        else if (n1 > n2)
            alpha = 1;
        else
            ok = false;
    }
    else if (rf_active) {
        alpha = 1;
        PRx = rf_ref_p[0], PRy = rf_ref_p[1];
    }
    else if (lf_active) {
        alpha = 0;
        PLx = lf_ref_p[0], PLy = lf_ref_p[1];
    }
    else
        ok = false;
    for (int k = 0; k < 6; ++k)
        left[k] = right[k] = 0;
    const double f_right = -alpha * mass * 9.81, f_left = -(1 - alpha) * mass * 9.81;
    const double t0x = -(PRx - zx) * f_right - (PLx - zx) * f_left, t0y = -(PRy - zy) * f_right - (PLy - zy) * f_left;
    double tRx, tLx;
    const double tRy = alpha * t0y, tLy = (1 - alpha) * t0y;
    if (t0x < 0)
        tRx = t0x, tLx = 0;
    else
        tRx = 0, tLx = t0x;
    left[2] = f_left, left[3] = tLy, left[4] = -tLx;
    right[2] = f_right, right[3] = tRy, right[4] = -tRx;
}

// stabilizer.cpp:168-224 on the SE3 sample `se3` (task) and `ct` (contact), both 24 doubles, in place
__device__ void ankle_admittance(double dt, const double* p, double copx, double copy, V3d foot, const double* se3_ref, const double* ct_ref, double* se3,
    double* ct)
{
    const double pitch = p[0] * (copx - foot.x), roll = p[1] * (copy - foot.y);
    double R[9], e[3];
    sample_rotation(se3_ref, R);
    euler012(R, e);
    e[0] += roll;
    e[1] += pitch;
    for (int k = 0; k < 3; ++k)
        se3[k] = se3_ref[k], ct[k] = ct_ref[k];
    rot_xyz_to_sample(e, se3);
    for (int k = 3; k < 12; ++k)
        ct[k] = se3[k];
    for (int k = 12; k < 24; ++k)
        se3[k] = se3_ref[k];
    se3[12 + 4] += p[2] * pitch / dt;
    se3[12 + 3] += p[3] * roll / dt;
    se3[18 + 4] += p[4] * pitch / (dt * dt);
    se3[18 + 3] += p[5] * roll / (dt * dt);
    for (int k = 12; k < 24; ++k)
        ct[k] = se3[k];
}

__global__ void k_stab_pre(const iwbc_problem_t* __restrict__ P, StabIO s)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.B)
        return;
    const int cap = s.cap, nref = s.nref;
    const double dt = P->dt;
    const double* rin = s.refs + (size_t)i * s.refs_stride;
    double* r = s.out + (size_t)i * nref;
    for (int k = 0; k < nref; ++k)
        r[k] = rin[k];
    const unsigned mask = s.mask ? s.mask[i] : s.mask_all;
    const int cl = P->stab_contact_lf, cr = P->stab_contact_rf;
    const bool lf_active = (mask >> cl) & 1u, rf_active = (mask >> cr) & 1u;
    // ---- contact-mask bucket of this robot (every launch set of the cycle is shape-uniform)
    if (s.blist) {
        const int k = mask == s.m0 ? 0 : (mask == s.m1 ? 1 : (mask == s.m2 ? 2 : -1));
        if (k >= 0)
            s.blist[(size_t)k * cap + atomicAdd(s.bcount + k, 1)] = i; // order inside a bucket is arbitrary; no result depends on it
        else if (s.status)
            s.status[i] = IWBC_STATUS_ERROR; // neither foot in contact: the reference's distributor throws ("is talos jumping ?")
    }
    // ---- set_behavior_type (humanoid_pos_tracker.cpp:116-132): every filter window takes the new parameter set's size
    const int bt = s.btype ? s.btype[i] : s.btype_all;
    const iwbc_stab_conf_t& conf = P->stab[bt];
    if (s.btype_state[i] != bt) {
        const int w_old = s.wsize[i], w_new = conf.filter_size;
        if (w_new != w_old) {
            for (int f = 0; f < IWBC_STAB_NF; ++f)
                filter_resize(s, i, f, w_old, w_new);
            s.wsize[i] = w_new;
        }
        s.btype_state[i] = bt;
    }
    const double* sens = s.sensors + (size_t)i * 15;
    const double *lf_force = sens, *rf_force = sens + 3, *lf_torque = sens + 6, *rf_torque = sens + 9, *imu_vel = sens + 12;
    // ---- what tsid's data holds: the state of the previous solve (or of the constructor)
    double com9[9];
    V3d lf_p, rf_p, imu_w;
    stab_model(P, s.qdata + (size_t)i * P->nq, s.vdata + (size_t)i * P->nv, com9, lf_p, rf_p, imu_w);
    const int activated = P->stabilizer_activated;
    // ---- TalosPosTracker::update: IMU momentum admittance (talos_pos_tracker.cpp:135-146; stabilizer.cpp:107-140)
    if (P->stab_talos && activated && conf.use_momentum && P->stab_task_momentum >= 0) {
        double fl[3];
        filter_push(s, i, 7, 3, imu_vel, fl);
        // (the call hands the MODEL's angular velocity in as `imu_angular_vel` and the filtered IMU as `model_angular_vel`)
        const double e0 = fl[0] - imu_w.x, e1 = fl[1] - imu_w.y, e2 = fl[2] - imu_w.z;
        const double vec[6] = {e0, e1, e2, e1, e0, e2};
        double* m = r + P->tasks[P->stab_task_momentum].ref_off;
        for (int k = 0; k < 6; ++k) {
            double vel = m[6 + k] - conf.momentum_p[k] * vec[k], acc = m[12 + k] - conf.momentum_d[k] * vec[k];
            vel = fmin(fmax(vel, -5.0), 5.0); // max_ref = 5 (stabilizer.hpp:38)
            acc = fmin(fmax(acc, -5.0), 5.0);
            m[6 + k] = vel;
            m[12 + k] = acc / dt;
        }
    }
    // ---- Cop::update (cop.cpp:10-58): per-foot CoP while the foot is loaded, total CoP when both are
    bool have_cop = false, have_l = false, have_r = false;
    double cop[2], lcop[2], rcop[2];
    {
        double lraw[2], rraw[2];
        const bool l_on = lf_force[2] > 30.0, r_on = rf_force[2] > 30.0; // estimators::FMIN
        if (l_on) {
            lraw[0] = lf_p.x + (-lf_torque[1] - lf_p.z * lf_force[0]) / lf_force[2];
            lraw[1] = lf_p.y + (lf_torque[0] - lf_p.z * lf_force[1]) / lf_force[2];
            filter_push(s, i, 1, 2, lraw, lcop);
        }
        else
            filter_reset(s, i, 1);
        if (r_on) {
            rraw[0] = rf_p.x + (-rf_torque[1] - rf_p.z * rf_force[0]) / rf_force[2];
            rraw[1] = rf_p.y + (rf_torque[0] - rf_p.z * rf_force[1]) / rf_force[2];
            filter_push(s, i, 2, 2, rraw, rcop);
        }
        else
            filter_reset(s, i, 2);
        bool fresh = false;
        if (l_on && r_on) {
            const double rl = lf_force[2] / (lf_force[2] + rf_force[2]), rr = 1 - rl;
            double raw[2];
            const double lfp[2] = {lf_p.x, lf_p.y}, rfp[2] = {rf_p.x, rf_p.y};
            for (int k = 0; k < 2; ++k) {
                const double in_l = rl * lraw[k] + rr * (rraw[k] + rfp[k] - lfp[k]);
                const double in_r = rl * (lraw[k] + lfp[k] - rfp[k]) + rr * rraw[k];
                raw[k] = in_l * rl + in_r * rr;
            }
            filter_push(s, i, 0, 2, raw, cop);
            fresh = true;
        }
        const int w = s.wsize[i];
        have_cop = fresh && s.cnt[(size_t)0 * cap + i] >= w && !(isnan(cop[0]) || isnan(cop[1]));
        have_l = l_on && s.cnt[(size_t)1 * cap + i] >= w && !(isnan(lcop[0]) || isnan(lcop[1]));
        have_r = r_on && s.cnt[(size_t)2 * cap + i] >= w && !(isnan(rcop[0]) || isnan(rcop[1]));
    }
    if (!activated)
        return;
    // ---- force filters, only while the foot is loaded (humanoid_pos_tracker.cpp:187-203)
    double lff[3], rff[3], tmp[3];
    if (sqrt(lf_force[0] * lf_force[0] + lf_force[1] * lf_force[1] + lf_force[2] * lf_force[2]) > 30.0) {
        filter_push(s, i, 3, 3, lf_force, lff);
        filter_push(s, i, 5, 3, lf_torque, tmp);
    }
    else
        lff[0] = lff[1] = lff[2] = 0;
    if (sqrt(rf_force[0] * rf_force[0] + rf_force[1] * rf_force[1] + rf_force[2] * rf_force[2]) > 30.0) {
        filter_push(s, i, 4, 3, rf_force, rff);
        filter_push(s, i, 6, 3, rf_torque, tmp);
    }
    else
        rff[0] = rff[1] = rff[2] = 0;
    for (int k = 0; k < 3; ++k) {
        s.ffilt[(size_t)k * cap + i] = lff[k];
        s.ffilt[(size_t)(3 + k) * cap + i] = rff[k];
    }
    const double* valid = have_cop ? cop : (have_l ? lcop : (have_r ? rcop : nullptr)); // :209
    const int ocl = P->contacts[cl].ref_off, ocr = P->contacts[cr].ref_off;
    // the contacts' force references as the contact objects hold them (:150-152)
    for (int k = 0; k < 6; ++k) {
        r[ocl + IWBC_REF_CONTACT_FREF + k] = s.base_fref[(size_t)k * cap + i];
        r[ocr + IWBC_REF_CONTACT_FREF + k] = s.base_fref[(size_t)(6 + k) * cap + i];
    }
    const double zmp_ref_x = com9[0] - com9[2] / 9.81 * com9[6], zmp_ref_y = com9[1] - com9[2] / 9.81 * com9[7]; // com_to_zmp (:9-20)
    // ---- com_admittance (:211-214; stabilizer.cpp:31-59)
    if (valid) {
        double* c = r + P->tasks[P->stab_task_com].ref_off;
        const double corx = zmp_ref_x - valid[0], cory = zmp_ref_y - valid[1];
        const double* g = conf.com_gains;
        c[0] -= g[0] * corx, c[1] -= g[1] * cory;
        c[3] -= g[2] * corx / dt, c[4] -= g[3] * cory / dt;
        c[6] -= g[4] * corx / (dt * dt), c[7] -= g[5] * cory / (dt * dt);
    }
    // ---- ZMP distributor admittance (:217-226; stabilizer.cpp:277-311)
    int wreg = s.wreg[i];
    if (valid && conf.use_zmp) {
        double lt[6], rt[6], ls[6], rs[6];
        bool ok1, ok2;
        zmp_distributor(P->total_mass, zmp_ref_x, zmp_ref_y, rin + ocl, rin + ocr, lf_active, rf_active, lt, rt, ok1);
        zmp_distributor(P->total_mass, valid[0], valid[1], rin + ocl, rin + ocr, lf_active, rf_active, ls, rs, ok2);
        if (ok1 && ok2) {
            for (int k = 0; k < 6; ++k) {
                r[ocl + IWBC_REF_CONTACT_FREF + k] = lt[k] - conf.zmp_p[k] * (lt[k] - ls[k]) - conf.zmp_d[k] * (lt[k] - ls[k]) / dt;
                r[ocr + IWBC_REF_CONTACT_FREF + k] = rt[k] - conf.zmp_p[k] * (rt[k] - rs[k]) - conf.zmp_d[k] * (rt[k] - rs[k]) / dt;
            }
            wreg = conf.zmp_w_variant;
            s.wreg[i] = wreg;
        }
    }
    r[ocl + IWBC_REF_CONTACT_WREG] = (double)wreg; // never put back by the reference (:291-301 restore references only)
    r[ocr + IWBC_REF_CONTACT_WREG] = (double)wreg;
    // ... and the force reference only of the ACTIVATED contacts (:297-300): a lifted foot keeps what the distributor wrote
    for (int k = 0; k < 6; ++k) {
        if (!lf_active)
            s.base_fref[(size_t)k * cap + i] = r[ocl + IWBC_REF_CONTACT_FREF + k];
        if (!rf_active)
            s.base_fref[(size_t)(6 + k) * cap + i] = r[ocr + IWBC_REF_CONTACT_FREF + k];
    }
    // ---- ankle admittances (:229-244)
    if (have_l && lf_active) {
        const int o = P->tasks[P->stab_task_lf].ref_off;
        ankle_admittance(dt, conf.ankle_gains, lcop[0], lcop[1], lf_p, rin + o, rin + ocl, r + o, r + ocl);
    }
    if (have_r && rf_active) {
        const int o = P->tasks[P->stab_task_rf].ref_off;
        ankle_admittance(dt, conf.ankle_gains, rcop[0], rcop[1], rf_p, rin + o, rin + ocr, r + o, r + ocr);
    }
    // ---- foot force difference admittance (:247-261; stabilizer.cpp:226-275): normal forces of the PREVIOUS solve
    const int w = s.wsize[i];
    const unsigned both = (1u << cl) | (1u << cr);
    if (P->stab_task_torso >= 0 && ((unsigned)s.have_prev[i] & both) == both && s.cnt[(size_t)3 * cap + i] >= w && s.cnt[(size_t)4 * cap + i] >= w) {
        const double zctrl = fabs(s.fn_prev[i] - s.fn_prev[(size_t)cap + i]) - fabs(lff[2] - rff[2]);
        const double torso_roll = conf.ffda_gains[0] * zctrl;
        const int o = P->tasks[P->stab_task_torso].ref_off;
        double R[9], e[3];
        sample_rotation(rin + o, R);
        euler012(R, e);
        e[0] += torso_roll;
        rot_xyz_to_sample(e, r + o);
        r[o + 12 + 3] += conf.ffda_gains[1] * torso_roll / dt;
        r[o + 18 + 3] += conf.ffda_gains[2] * torso_roll / (dt * dt);
    }
}

// after the solve: what the next tick's pre-step reads
__global__ void k_stab_post(const iwbc_problem_t* __restrict__ P, int B, int cap, const double* __restrict__ q, const double* __restrict__ v,
    const unsigned* __restrict__ mask, unsigned mask_all, const int* __restrict__ status, const double* __restrict__ f, double* __restrict__ qdata,
    double* __restrict__ vdata, double* __restrict__ fn_prev, int* __restrict__ have_prev)
{
    const int nq = P->nq, nv = P->nv, w = nq + nv;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < (long)B * w; e += (long)gridDim.x * blockDim.x) {
        const int i = (int)(e / w), k = (int)(e - (long)i * w);
        // computeProblemData ran on (q, v) whatever the solver said afterwards
        if (k < nq)
            qdata[(size_t)i * nq + k] = q[(size_t)i * nq + k];
        else
            vdata[(size_t)i * nv + (k - nq)] = v[(size_t)i * nv + (k - nq)];
        if (k == 0 && status[i] == IWBC_STATUS_OPTIMAL) { // activated_contacts_forces_ (controller.cpp:258-261); a failed solve leaves it alone
            const unsigned m = mask ? mask[i] : mask_all;
            const double* fi = f + (size_t)i * 12 * P->ncontacts;
            const int cs[2] = {P->stab_contact_lf, P->stab_contact_rf};
            for (int s2 = 0; s2 < 2; ++s2) { // Contact6d::getNormalForce: sum over the four corners of n . f_i (row 16 of the force block = [n n n n])
                const double* fc = fi + 12 * cs[s2];
                const double* nrow = P->contacts[cs[s2]].B + 16 * 12;
                double fn = 0;
                for (int c = 0; c < 12; ++c)
                    fn += nrow[c] * fc[c];
                fn_prev[(size_t)s2 * cap + i] = fn;
            }
            have_prev[i] = (int)m;
        }
    }
}

__global__ void k_stab_reset(const iwbc_problem_t* __restrict__ P, int cap, int W, double* buf, int* cnt, int* wsize, int* btype, double* ffilt, double* fn_prev,
    int* have_prev, int* wreg, double* base_fref, double* qdata, double* vdata)
{
    const long n = (long)IWBC_STAB_NF * 3 * W * cap;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
        buf[e] = 0.0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < cap; i += (long)gridDim.x * blockDim.x) {
        for (int f = 0; f < IWBC_STAB_NF; ++f)
            cnt[(size_t)f * cap + i] = 0;
        wsize[i] = P->stab[IWBC_BT_FIXED_BASE].filter_size; // parse_stabilizer sizes everything for FIXED_BASE (:100-113)
        btype[i] = IWBC_BT_FIXED_BASE;
        for (int k = 0; k < 6; ++k)
            ffilt[(size_t)k * cap + i] = 0.0;
        fn_prev[i] = fn_prev[(size_t)cap + i] = 0.0;
        have_prev[i] = 0;
        wreg[i] = 0;
        for (int k = 0; k < 12; ++k)
            base_fref[(size_t)k * cap + i] = 0.0;
        for (int k = 0; k < P->nq; ++k)
            qdata[(size_t)i * P->nq + k] = P->q0[k]; // the constructor's computeProblemData (pos_tracker.cpp:105-106)
        for (int k = 0; k < P->nv; ++k)
            vdata[(size_t)i * P->nv + k] = 0.0;
    }
}

int ensure_problem_on_device(iwbc_batch* b, cudaStream_t st)
{
    if (b->dprob_revision != b->prob->revision) {
        CUDA_TRY(cudaMemcpyAsync(b->d_prob, &b->prob->cs.prob, sizeof(iwbc_problem_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        b->dprob_revision = b->prob->revision;
    }
    return IWBC_OK;
}

} // namespace

extern "C" {

iwbc_stabilizer* iwbc_stabilizer_create(iwbc_batch* b)
{
    if (!b) {
        set_error("iwbc_stabilizer_create: NULL batch");
        return nullptr;
    }
    const iwbc_problem_t& P = b->prob->cs.prob;
    if (!P.stab_present) {
        set_error("iwbc_stabilizer_create: this controller has no stabilizer parameter sets (humanoid-pos-tracker / talos-pos-tracker only)");
        return nullptr;
    }
    if (cudaSetDevice(b->device) != cudaSuccess) {
        set_error("iwbc_stabilizer_create: cannot select the batch's device");
        return nullptr;
    }
    auto* s = new iwbc_stabilizer;
    s->batch = b, s->device = b->device, s->cap = b->max_batch;
    s->W = std::max(P.stab[0].filter_size, std::max(P.stab[1].filter_size, P.stab[2].filter_size));
    const size_t cap = s->cap;
    bool ok = cudaMalloc((void**)&s->d_buf, (size_t)IWBC_STAB_NF * 3 * s->W * cap * 8) == cudaSuccess && cudaMalloc((void**)&s->d_cnt, IWBC_STAB_NF * cap * 4) == cudaSuccess
        && cudaMalloc((void**)&s->d_wsize, cap * 4) == cudaSuccess && cudaMalloc((void**)&s->d_btype, cap * 4) == cudaSuccess
        && cudaMalloc((void**)&s->d_ffilt, 6 * cap * 8) == cudaSuccess && cudaMalloc((void**)&s->d_fn_prev, 2 * cap * 8) == cudaSuccess
        && cudaMalloc((void**)&s->d_have_prev, cap * 4) == cudaSuccess && cudaMalloc((void**)&s->d_wreg, cap * 4) == cudaSuccess
        && cudaMalloc((void**)&s->d_base_fref, 12 * cap * 8) == cudaSuccess && cudaMalloc((void**)&s->d_qdata, cap * P.nq * 8) == cudaSuccess
        && cudaMalloc((void**)&s->d_vdata, cap * P.nv * 8) == cudaSuccess && cudaMalloc((void**)&s->d_refs, cap * P.nref * 8) == cudaSuccess
        && cudaMalloc((void**)&s->d_blist, 3 * cap * 4) == cudaSuccess && cudaMalloc((void**)&s->d_bcount, 4 * 4) == cudaSuccess;
    s->masks[0] = (1u << P.stab_contact_lf) | (1u << P.stab_contact_rf);
    s->masks[1] = 1u << P.stab_contact_lf;
    s->masks[2] = 1u << P.stab_contact_rf;
    if (!ok || iwbc_stabilizer_reset(s, nullptr) != IWBC_OK) {
        if (std::string(iwbc_last_error()).empty())
            set_error("iwbc_stabilizer_create: CUDA allocation failed");
        iwbc_stabilizer_free(s);
        return nullptr;
    }
    return s;
}

void iwbc_stabilizer_free(iwbc_stabilizer* s)
{
    if (!s)
        return;
    cudaSetDevice(s->device);
    void* ptrs[] = {s->d_buf, s->d_cnt, s->d_wsize, s->d_btype, s->d_ffilt, s->d_fn_prev, s->d_have_prev, s->d_wreg, s->d_base_fref, s->d_qdata, s->d_vdata,
        s->d_refs, s->d_blist, s->d_bcount};
    for (void* p : ptrs)
        if (p)
            cudaFree(p);
    delete s;
}

int iwbc_stabilizer_reset(iwbc_stabilizer* s, void* stream)
{
    if (!s) {
        set_error("iwbc_stabilizer_reset: NULL handle");
        return IWBC_ERR_ARG;
    }
    iwbc_batch* b = s->batch;
    CUDA_TRY(cudaSetDevice(s->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    int rc = ensure_problem_on_device(b, st);
    if (rc)
        return rc;
    k_stab_reset<<<1184, 256, 0, st>>>(b->d_prob, s->cap, s->W, s->d_buf, s->d_cnt, s->d_wsize, s->d_btype, s->d_ffilt, s->d_fn_prev, s->d_have_prev, s->d_wreg,
        s->d_base_fref, s->d_qdata, s->d_vdata);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(st));
    return IWBC_OK;
}

int iwbc_batch_update_sensors_device(iwbc_batch* b, iwbc_stabilizer* s, int B, const double* q, const double* v, const double* refs, long refs_stride,
    const unsigned* contact_mask, const int* behavior_type, int behavior_type_all, const double* sensors, const iwbc_results_t* out, double* refs_used,
    void* stream)
{
    if (!b || !s || !q || !v || !refs || !sensors || !out || B < 0 || s->batch != b) {
        set_error("iwbc_batch_update_sensors_device: bad argument");
        return IWBC_ERR_ARG;
    }
    if (B > b->max_batch) {
        set_error("iwbc_batch_update_sensors_device: B exceeds max_batch");
        return IWBC_ERR_ARG;
    }
    if (!behavior_type && (behavior_type_all < 0 || behavior_type_all > 2)) {
        set_error("behavior type is either FIXED_BASE (0), SINGLE_SUPPORT (1) or DOUBLE_SUPPORT (2)"); // controller.cpp:438-439
        return IWBC_ERR_ARG;
    }
    if (B == 0)
        return IWBC_OK;
    const iwbc_problem_t& P = b->prob->cs.prob;
    if (refs_stride != 0 && refs_stride != P.nref) {
        set_error("iwbc_batch_update_sensors_device: refs_stride must be 0 (shared) or nref");
        return IWBC_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    int rc = ensure_problem_on_device(b, st);
    if (rc)
        return rc;
    double* rows = refs_used ? refs_used : s->d_refs;
    int* status = out->status ? out->status : b->d_status;
    double* f = out->f ? out->f : b->d_f;
    StabIO io;
    std::memset(&io, 0, sizeof(io));
    io.B = B, io.cap = s->cap, io.W = s->W, io.nref = P.nref, io.refs_stride = refs_stride, io.refs = refs;
    io.mask = contact_mask, io.mask_all = iwbc_problem_default_contact_mask(b->prob), io.btype = behavior_type, io.btype_all = behavior_type_all;
    io.sensors = sensors, io.buf = s->d_buf, io.cnt = s->d_cnt, io.wsize = s->d_wsize, io.btype_state = s->d_btype, io.ffilt = s->d_ffilt;
    io.fn_prev = s->d_fn_prev, io.have_prev = s->d_have_prev, io.wreg = s->d_wreg, io.base_fref = s->d_base_fref;
    io.qdata = s->d_qdata, io.vdata = s->d_vdata, io.out = rows;
    io.m0 = s->masks[0], io.m1 = s->masks[1], io.m2 = s->masks[2];
    io.status = status;
    if (contact_mask) {
        io.blist = s->d_blist, io.bcount = s->d_bcount;
        CUDA_TRY(cudaMemsetAsync(s->d_bcount, 0, 4 * sizeof(int), st));
    }
    k_stab_pre<<<(B + 63) / 64, 64, 0, st>>>(b->d_prob, io);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    b->sensor_call = true; // (the plain entry points refuse a controller whose stabilizer is activated: it needs the sensors)
    for (int k = 0; k < (contact_mask ? 3 : 1); ++k) {
        dev::IO cio;
        std::memset(&cio, 0, sizeof(cio));
        cio.B = B;
        cio.q = q, cio.v = v, cio.refs = rows, cio.refs_stride = P.nref;
        cio.tau = out->tau, cio.dv = out->dv, cio.f = f, cio.q_next = out->q_next, cio.v_next = out->v_next, cio.q_cmd = out->q_cmd;
        cio.momentum = out->momentum, cio.status = status, cio.iters = out->iters, cio.active = out->active;
        if (contact_mask)
            cio.in_list = s->d_blist + (size_t)k * s->cap, cio.in_count = s->d_bcount + k;
        rc = iwbc_run_cycle(b, cio, contact_mask ? s->masks[k] : io.mask_all, st);
        if (rc)
            break;
    }
    b->sensor_call = false;
    if (rc)
        return rc;
    k_stab_post<<<std::min(1184, (B * (P.nq + P.nv) + 255) / 256), 256, 0, st>>>(b->d_prob, B, s->cap, q, v, contact_mask, io.mask_all, status, f, s->d_qdata,
        s->d_vdata, s->d_fn_prev, s->d_have_prev);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    return IWBC_OK;
}

int iwbc_batch_update_sensors(iwbc_batch* b, iwbc_stabilizer* s, int B, const double* q, const double* v, const double* refs, long refs_stride,
    const unsigned* contact_mask, const int* behavior_type, int behavior_type_all, const double* sensors, const iwbc_results_t* out, double* refs_used)
{
    if (!b || !s || !q || !v || !refs || !sensors || !out || B < 0) {
        set_error("iwbc_batch_update_sensors: bad argument");
        return IWBC_ERR_ARG;
    }
    if (B > b->max_batch) {
        set_error("iwbc_batch_update_sensors: B exceeds max_batch");
        return IWBC_ERR_ARG;
    }
    if (B == 0)
        return IWBC_OK;
    CUDA_TRY(cudaSetDevice(b->device));
    const iwbc_problem_t& P = b->prob->cs.prob;
    cudaStream_t st = b->stream;
    const size_t nq = P.nq, nv = P.nv, na = P.na, nf = 12 * P.ncontacts, nref = P.nref;
    if (!b->d_phase) { // (shared with the behaviour roll-out: per-robot ints and masks)
        CUDA_TRY(cudaMalloc((void**)&b->d_phase, (size_t)b->max_batch * sizeof(int)));
        CUDA_TRY(cudaMalloc((void**)&b->d_cmask, (size_t)b->max_batch * sizeof(unsigned)));
    }
    if (!b->d_sensors)
        CUDA_TRY(cudaMalloc((void**)&b->d_sensors, (size_t)b->max_batch * 15 * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(b->d_q, q, B * nq * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_v, v, B * nv * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_refs, refs, (refs_stride ? (size_t)B : 1) * nref * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_sensors, sensors, (size_t)B * 15 * 8, cudaMemcpyHostToDevice, st));
    if (contact_mask)
        CUDA_TRY(cudaMemcpyAsync(b->d_cmask, contact_mask, B * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    if (behavior_type)
        CUDA_TRY(cudaMemcpyAsync(b->d_phase, behavior_type, B * sizeof(int), cudaMemcpyHostToDevice, st));
    iwbc_results_t r;
    r.tau = out->tau ? b->d_tau : nullptr, r.dv = out->dv ? b->d_dv : nullptr, r.f = b->d_f, r.q_next = out->q_next ? b->d_qn : nullptr;
    r.v_next = out->v_next ? b->d_vn : nullptr, r.q_cmd = out->q_cmd ? b->d_qc : nullptr, r.momentum = out->momentum ? b->d_mom : nullptr;
    r.status = b->d_status, r.iters = out->iters ? b->d_iters : nullptr, r.active = out->active ? b->d_active : nullptr;
    int rc = iwbc_batch_update_sensors_device(b, s, B, b->d_q, b->d_v, b->d_refs, refs_stride, contact_mask ? b->d_cmask : nullptr,
        behavior_type ? b->d_phase : nullptr, behavior_type_all, b->d_sensors, &r, nullptr, st);
    if (rc)
        return rc;
    auto back = [&](void* host, const void* devp, size_t bytes) -> int {
        if (host)
            CUDA_TRY(cudaMemcpyAsync(host, devp, bytes, cudaMemcpyDeviceToHost, st));
        return IWBC_OK;
    };
    rc = back(out->tau, b->d_tau, B * na * 8) | back(out->dv, b->d_dv, B * nv * 8) | back(out->f, b->d_f, B * nf * 8) | back(out->q_next, b->d_qn, B * nq * 8)
        | back(out->v_next, b->d_vn, B * nv * 8) | back(out->q_cmd, b->d_qc, B * nv * 8) | back(out->momentum, b->d_mom, B * 3 * 8)
        | back(out->status, b->d_status, B * 4) | back(out->iters, b->d_iters, B * 4) | back(out->active, b->d_active, (size_t)B * IWBC_ACTIVE_WORDS * 4)
        | back(refs_used, s->d_refs, B * nref * 8);
    if (rc)
        return IWBC_ERR_CUDA;
    CUDA_TRY(cudaStreamSynchronize(st));
    return IWBC_OK;
}

} // extern "C"
