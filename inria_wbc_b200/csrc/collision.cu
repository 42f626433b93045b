// collision.cu -- post-solve model collision check.
//
// The reference covers every listed link with spheres and, after each solve, tests all pairs of spheres that belong to DIFFERENT
// members against each other on the model pose (src/safety/collision_check.cpp:27-87, called at
// src/controllers/controller.cpp:309-310; config keys check_model_collisions / collision_path, controller.cpp:89-98).  Sphere
// centres ride on the placement of the joint that carries the link (data.oMi[frame.parent], collision_check.cpp:78); the fourth
// number of a sphere is a DIAMETER (two spheres collide when |c1 - c2| < d1/2 + d2/2, collision_check.cpp:40).
// Batched form: one thread per robot state recomputes the joint placements (forward kinematics over the flattened tree) and
// walks the member pairs in the reference's order (members sorted by name: it iterates a std::map), reporting the first hit.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "batch_state.hpp"
#include "yaml_lite.hpp"

using namespace iwbc;

struct iwbc_collision_sphere_t {
    double p[3];
    double radius; // half of the diameter the file gives
    int joint;     // supporting joint of the link (-1: universe)
    int member;
};

struct iwbc_collision {
    const iwbc_problem* prob = nullptr;
    std::vector<std::string> members;           // sorted by name
    std::vector<int> member_start;              // [nmembers + 1] into spheres
    std::vector<iwbc_collision_sphere_t> spheres; // grouped by member, file order inside a member
    int skipped_links = 0;
    int device = -1;
    iwbc_collision_sphere_t* d_spheres = nullptr;
    int* d_member_start = nullptr;
};

namespace {

#define IWBC_COL_MAXS 320

__device__ __forceinline__ void se3_mul(const double* A, const double* B, double* C)
{ // A, B, C: R row-major (9) then p (3);  C = A * B
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
        C[9 + i] = A[3 * i] * B[9] + A[3 * i + 1] * B[10] + A[3 * i + 2] * B[11] + A[9 + i];
    }
}

// ---- send_cmd latch / qp_step_back (controller.cpp:238-242, 263-282, 309-312, 445-450), one thread per (instance, entry)
// before the solve: instances that still send commands remember the state the tick starts from
__global__ void k_safety_pre(int B, int nq, int nv, const int* __restrict__ send_cmd, const double* __restrict__ q, const double* __restrict__ v,
    double* __restrict__ q_prev, double* __restrict__ v_prev)
{
    const int w = nq + nv;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < (long)B * w; e += (long)gridDim.x * blockDim.x) {
        const int i = (int)(e / w), k = (int)(e - (long)i * w);
        if (!send_cmd[i])
            continue;
        if (k < nq)
            q_prev[(size_t)i * nq + k] = q[(size_t)i * nq + k];
        else
            v_prev[(size_t)i * nv + (k - nq)] = v[(size_t)i * nv + (k - nq)];
    }
}
// after the solve (one warp per instance): q_solver_ always follows an OPTIMAL solve, the command vectors only while send_cmd_;
// then a model collision latches send_cmd_ = false unless the command is still all zero (!q_.isZero(1e-3))
__global__ void k_safety_post(int B, int nv, int na, const int* __restrict__ status, const int* __restrict__ colliding, const double* __restrict__ qc_new,
    const double* __restrict__ tau_new, const double* __restrict__ v_new, const double* __restrict__ dv_new, int* __restrict__ send_cmd,
    double* __restrict__ q_solver, double* __restrict__ q_cmd, double* __restrict__ tau_cmd, double* __restrict__ dq_cmd, double* __restrict__ ddq_cmd)
{
    const int lane = threadIdx.x & 31;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < B; i += (gridDim.x * blockDim.x) >> 5) {
        if (status[i] != IWBC_STATUS_OPTIMAL)
            continue; // the reference throws before any of this (controller.cpp:284-308)
        const int send = send_cmd[i];
        bool nonzero = false;
        for (int k = lane; k < nv; k += 32) {
            const double qn = qc_new[(size_t)i * nv + k];
            if (q_solver)
                q_solver[(size_t)i * nv + k] = qn;
            if (send) {
                q_cmd[(size_t)i * nv + k] = qn;
                if (dq_cmd)
                    dq_cmd[(size_t)i * nv + k] = v_new[(size_t)i * nv + k];
                if (ddq_cmd)
                    ddq_cmd[(size_t)i * nv + k] = dv_new[(size_t)i * nv + k];
            }
            nonzero = nonzero || fabs(q_cmd[(size_t)i * nv + k]) > 1e-3;
        }
        if (send && tau_cmd)
            for (int k = lane; k < na; k += 32)
                tau_cmd[(size_t)i * na + k] = tau_new[(size_t)i * na + k];
        nonzero = __any_sync(0xffffffffu, nonzero);
        if (lane == 0 && colliding && colliding[i] && nonzero)
            send_cmd[i] = 0;
    }
}
__global__ void k_step_back(int B, int nq, int nv, const int* __restrict__ select, const double* __restrict__ q_prev, const double* __restrict__ v_prev,
    double* __restrict__ q, double* __restrict__ v)
{
    const int w = nq + nv;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < (long)B * w; e += (long)gridDim.x * blockDim.x) {
        const int i = (int)(e / w), k = (int)(e - (long)i * w);
        if (select && !select[i])
            continue;
        if (k < nq)
            q[(size_t)i * nq + k] = q_prev[(size_t)i * nq + k];
        else
            v[(size_t)i * nv + (k - nq)] = v_prev[(size_t)i * nv + (k - nq)];
    }
}

// one thread per state
__global__ void k_collision(const iwbc_problem_t* __restrict__ P, const iwbc_collision_sphere_t* __restrict__ S, const int* __restrict__ mstart,
    int nmembers, int nspheres, int B, const double* __restrict__ q, int* __restrict__ colliding, int* __restrict__ pair4)
{
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= B)
        return;
    const double* qi = q + (size_t)inst * P->nq;
    double oM[IWBC_MAXJ][12];
    for (int j = 0; j < P->nj; ++j) { // joints are DFS-numbered: a parent precedes its children
        double jm[12];
        const int ty = P->jtype[j];
        const double* qq = qi + P->idx_q[j];
        if (ty == IWBC_J_FREEFLYER) {
            const double x = qq[3], y = qq[4], z = qq[5], w = qq[6];
            jm[0] = 1 - 2 * (y * y + z * z), jm[1] = 2 * (x * y - z * w), jm[2] = 2 * (x * z + y * w);
            jm[3] = 2 * (x * y + z * w), jm[4] = 1 - 2 * (x * x + z * z), jm[5] = 2 * (y * z - x * w);
            jm[6] = 2 * (x * z - y * w), jm[7] = 2 * (y * z + x * w), jm[8] = 1 - 2 * (x * x + y * y);
            jm[9] = qq[0], jm[10] = qq[1], jm[11] = qq[2];
        }
        else if (ty == IWBC_J_REVOLUTE) {
            const double ax = P->axis[j][0], ay = P->axis[j][1], az = P->axis[j][2];
            double s, c;
            sincos(qq[0], &s, &c);
            const double oc = 1 - c;
            jm[0] = c + ax * ax * oc, jm[1] = ax * ay * oc - az * s, jm[2] = ax * az * oc + ay * s;
            jm[3] = ay * ax * oc + az * s, jm[4] = c + ay * ay * oc, jm[5] = ay * az * oc - ax * s;
            jm[6] = az * ax * oc - ay * s, jm[7] = az * ay * oc + ax * s, jm[8] = c + az * az * oc;
            jm[9] = jm[10] = jm[11] = 0;
        }
        else { // prismatic
            jm[0] = jm[4] = jm[8] = 1;
            jm[1] = jm[2] = jm[3] = jm[5] = jm[6] = jm[7] = 0;
            jm[9] = qq[0] * P->axis[j][0], jm[10] = qq[0] * P->axis[j][1], jm[11] = qq[0] * P->axis[j][2];
        }
        double li[12];
        se3_mul(P->jplace[j], jm, li);
        const int pa = P->parent[j];
        if (pa >= 0)
            se3_mul(oM[pa], li, oM[j]);
        else
            for (int k = 0; k < 12; ++k)
                oM[j][k] = li[k];
    }
    double c[IWBC_COL_MAXS][3];
    for (int s = 0; s < nspheres; ++s) {
        const iwbc_collision_sphere_t sp = S[s];
        if (sp.joint < 0) {
            c[s][0] = sp.p[0], c[s][1] = sp.p[1], c[s][2] = sp.p[2];
            continue;
        }
        const double* M = oM[sp.joint];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            c[s][k] = M[3 * k] * sp.p[0] + M[3 * k + 1] * sp.p[1] + M[3 * k + 2] * sp.p[2] + M[9 + k];
    }
    int hit = 0, ha = -1, hi = -1, hb = -1, hj = -1;
    for (int a = 0; a < nmembers && !hit; ++a)
        for (int b2 = 0; b2 < nmembers && !hit; ++b2) {
            if (a == b2)
                continue;
            for (int i = mstart[a]; i < mstart[a + 1] && !hit; ++i) {
                const double ri = S[i].radius;
                for (int j = mstart[b2]; j < mstart[b2 + 1]; ++j) {
                    const double dx = c[j][0] - c[i][0], dy = c[j][1] - c[i][1], dz = c[j][2] - c[i][2];
                    if (sqrt(dx * dx + dy * dy + dz * dz) < S[j].radius + ri) {
                        hit = 1, ha = a, hi = i - mstart[a], hb = b2, hj = j - mstart[b2];
                        break;
                    }
                }
            }
        }
    colliding[inst] = hit;
    if (pair4) {
        int* o = pair4 + 4 * (size_t)inst;
        o[0] = ha, o[1] = hi, o[2] = hb, o[3] = hj;
    }
}

} // namespace

extern "C" {

iwbc_collision* iwbc_collision_load(const iwbc_problem* p, const char* collision_yaml)
{
    if (!p || !collision_yaml) {
        set_error("iwbc_collision_load: NULL argument");
        return nullptr;
    }
    iwbc_collision* h = new iwbc_collision();
    h->prob = p;
    try {
        yaml::Node root = yaml::LoadFile(collision_yaml);
        const yaml::Node& mem = root["members"];
        if (!mem.is_map())
            throw host::ConfigError(std::string("collision file without a 'members' map: ") + collision_yaml);
        // the reference keeps the members in a std::map: iteration order = sorted by name (collision_check.hpp)
        std::map<std::string, const yaml::Node*> sorted;
        for (auto& kv : mem.map)
            sorted[kv.first] = &kv.second;
        const host::Model& M = p->cs.model;
        h->member_start.push_back(0);
        for (auto& kv : sorted) {
            h->members.push_back(kv.first);
            if (!kv.second->is_map())
                throw host::ConfigError("collisions yaml: member '" + kv.first + "' is not a map of links");
            for (auto& link : kv.second->map) {
                int fid = M.frame_id(link.first);
                if (fid < 0) { // the reference only visits frames the model has (collision_check.cpp:69-70)
                    h->skipped_links++;
                    continue;
                }
                if (!link.second.is_seq())
                    throw host::ConfigError("collisions yaml: link '" + link.first + "' needs a list of spheres");
                for (size_t s = 0; s < link.second.size(); ++s) {
                    const std::vector<double> v = link.second[s].as<std::vector<double>>();
                    if (v.size() != 4)
                        throw host::ConfigError("collisions yaml : sphere data should be an float array of dim 4"); // collision_check.cpp:73
                    iwbc_collision_sphere_t sp;
                    // (the reference reads the four numbers as float: collision_check.cpp:20,65)
                    sp.p[0] = (float)v[0], sp.p[1] = (float)v[1], sp.p[2] = (float)v[2];
                    sp.radius = (double)((float)v[3] / 2 + 0.0f);
                    sp.joint = M.frames[fid].parent;
                    sp.member = (int)h->members.size() - 1;
                    h->spheres.push_back(sp);
                }
            }
            h->member_start.push_back((int)h->spheres.size());
        }
        if (h->spheres.empty())
            throw host::ConfigError("collisions yaml: no sphere sits on a link of this model");
        if ((int)h->spheres.size() > IWBC_COL_MAXS)
            throw host::ConfigError("collisions yaml: more than " + std::to_string(IWBC_COL_MAXS) + " spheres");
    }
    catch (std::exception& e) {
        set_error(e.what());
        delete h;
        return nullptr;
    }
    return h;
}

void iwbc_collision_free(iwbc_collision* h)
{
    if (!h)
        return;
    if (h->d_spheres) {
        cudaSetDevice(h->device);
        cudaFree(h->d_spheres);
        cudaFree(h->d_member_start);
    }
    delete h;
}

int iwbc_collision_info(const iwbc_collision* h, int* nspheres, int* nmembers, int* skipped_links)
{
    if (!h)
        return IWBC_ERR_ARG;
    if (nspheres)
        *nspheres = (int)h->spheres.size();
    if (nmembers)
        *nmembers = (int)h->members.size();
    if (skipped_links)
        *skipped_links = h->skipped_links;
    return IWBC_OK;
}

int iwbc_collision_spheres(const iwbc_collision* h, double* p_radius4, int* joint_member2)
{
    if (!h)
        return IWBC_ERR_ARG;
    for (size_t s = 0; s < h->spheres.size(); ++s) {
        if (p_radius4) {
            for (int k = 0; k < 3; ++k)
                p_radius4[4 * s + k] = h->spheres[s].p[k];
            p_radius4[4 * s + 3] = h->spheres[s].radius;
        }
        if (joint_member2) {
            joint_member2[2 * s] = h->spheres[s].joint;
            joint_member2[2 * s + 1] = h->spheres[s].member;
        }
    }
    return IWBC_OK;
}

int iwbc_batch_check_collisions_device(iwbc_batch* b, iwbc_collision* h, int B, const double* q, int* colliding, int* pair4, void* stream)
{
    if (!b || !h || !q || !colliding || B < 0) {
        set_error("iwbc_batch_check_collisions_device: bad argument");
        return IWBC_ERR_ARG;
    }
    if (h->prob != b->prob) {
        set_error("iwbc_batch_check_collisions_device: the collision model was loaded for another controller");
        return IWBC_ERR_ARG;
    }
    if (B == 0)
        return IWBC_OK;
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    if (!h->d_spheres) {
        CUDA_TRY(cudaMalloc((void**)&h->d_spheres, h->spheres.size() * sizeof(iwbc_collision_sphere_t)));
        CUDA_TRY(cudaMalloc((void**)&h->d_member_start, h->member_start.size() * sizeof(int)));
        CUDA_TRY(cudaMemcpy(h->d_spheres, h->spheres.data(), h->spheres.size() * sizeof(iwbc_collision_sphere_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->d_member_start, h->member_start.data(), h->member_start.size() * sizeof(int), cudaMemcpyHostToDevice));
        h->device = b->device;
    }
    if (b->dprob_revision != b->prob->revision) { // (the cycle launches refresh the device copy of the problem too)
        CUDA_TRY(cudaMemcpyAsync(b->d_prob, &b->prob->cs.prob, sizeof(iwbc_problem_t), cudaMemcpyHostToDevice, st));
        b->dprob_revision = b->prob->revision;
    }
    k_collision<<<(B + 63) / 64, 64, 0, st>>>(b->d_prob, h->d_spheres, h->d_member_start, (int)h->members.size(), (int)h->spheres.size(), B, q, colliding,
        pair4);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    return IWBC_OK;
}

int iwbc_batch_safe_rollout_device(iwbc_batch* b, iwbc_collision* h, int B, double* q, double* v, const double* refs, long refs_stride,
    unsigned contact_mask, int ticks, const iwbc_safety_t* s, int* status, void* stream)
{
    if (!b || !q || !v || !refs || !s || !s->send_cmd || !s->q_prev || !s->v_prev || !s->q_cmd || B < 0 || ticks < 0 || (h && !s->colliding)) {
        set_error("iwbc_batch_safe_rollout_device: bad argument");
        return IWBC_ERR_ARG;
    }
    if (B > b->max_batch) {
        set_error("iwbc_batch_safe_rollout_device: B exceeds max_batch");
        return IWBC_ERR_ARG;
    }
    if (B == 0 || ticks == 0)
        return IWBC_OK;
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    const iwbc_problem_t& P = b->prob->cs.prob;
    const int grid = std::min(1184, (B * (P.nq + P.nv) + 255) / 256);
    for (int t = 0; t < ticks; ++t) {
        k_safety_pre<<<grid, 256, 0, st>>>(B, P.nq, P.nv, s->send_cmd, q, v, s->q_prev, s->v_prev);
        b->launches++;
        if (h) { // check_model_collisions: on the model data of this tick, i.e. the state the tick starts from (controller.cpp:309-310)
            int rc = iwbc_batch_check_collisions_device(b, h, B, q, s->colliding, nullptr, st);
            if (rc)
                return rc;
        }
        iwbc_results_t r;
        std::memset(&r, 0, sizeof(r));
        r.tau = b->d_tau, r.dv = b->d_dv, r.q_cmd = b->d_qc, r.status = status ? status : b->d_status;
        // one open-loop tick on the state in place; a QP that is not solved to optimality freezes the instance (roll-out semantics)
        int rc = iwbc_batch_rollout_device(b, B, q, v, refs, refs_stride, contact_mask, 1, &r, st);
        if (rc)
            return rc;
        k_safety_post<<<std::min(1184, (B + 7) / 8), 256, 0, st>>>(B, P.nv, P.na, r.status, h ? s->colliding : nullptr, b->d_qc, b->d_tau, v, b->d_dv,
            s->send_cmd, s->q_solver, s->q_cmd, s->tau_cmd, s->dq_cmd, s->ddq_cmd);
        b->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return IWBC_OK;
}

int iwbc_batch_qp_step_back_device(iwbc_batch* b, int B, double* q, double* v, const iwbc_safety_t* s, const int* select, void* stream)
{
    if (!b || !q || !v || !s || !s->q_prev || !s->v_prev || B < 0) {
        set_error("iwbc_batch_qp_step_back_device: bad argument");
        return IWBC_ERR_ARG;
    }
    if (B == 0)
        return IWBC_OK;
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    const iwbc_problem_t& P = b->prob->cs.prob;
    k_step_back<<<std::min(1184, (B * (P.nq + P.nv) + 255) / 256), 256, 0, st>>>(B, P.nq, P.nv, select, s->q_prev, s->v_prev, q, v);
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    return IWBC_OK;
}

} // extern "C"
