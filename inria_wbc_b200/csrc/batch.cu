// batch.cu -- GPU half of the C ABI (include/iwbc_batch.h): kernel entry points, launch logic,
// device workspaces, host<->device staging.  sm_100a only; no CPU fallback.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/iwbc_batch.h"
#include "cycle_kernel.cuh"
#include "fast_kernel.cuh"
#include "host_state.hpp"

namespace iwbc {
namespace dev {

    // FP64 FMA peak probe: 8 independent dependent-FMA chains per thread (the roofline denominator of this
    // path is FP64 CUDA-core throughput, which MEASURED_PEAKS.json does not hold: SURVEY.md 8(d))
    __global__ void k_fp64_peak(double* out, int iters, double a, double b)
    {
        double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
        for (int i = 0; i < iters; ++i) {
            x0 = fma(x0, a, b), x1 = fma(x1, a, b), x2 = fma(x2, a, b), x3 = fma(x3, a, b);
            x4 = fma(x4, a, b), x5 = fma(x5, a, b), x6 = fma(x6, a, b), x7 = fma(x7, a, b);
        }
        if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 123.456)
            out[0] = x0;
    }

} // namespace dev
} // namespace iwbc

using namespace iwbc;

#include "batch_state.hpp"

namespace {

std::mutex g_attr_mutex;

// fast pass + active-set continuation: one translation unit per compile-time (nv, me) shape (fast_tu.cuh)
int launch_fast(iwbc_batch* b, const dev::IO& io, cudaStream_t st)
{
    const int nv = b->prob->cs.prob.nv, me = b->fshape.me;
    if (nv == 50 && me == 18)
        return iwbc_launch_fast_50_18(b, io, st); // Talos (full URDF), double support
    if (nv == 50 && me == 12)
        return iwbc_launch_fast_50_12(b, io, st); // Talos, single support
    if (nv == 38 && me == 18)
        return iwbc_launch_fast_38_18(b, io, st); // iCub / Talos fast URDF, double support
    if (nv == 38 && me == 12)
        return iwbc_launch_fast_38_12(b, io, st);
    return iwbc_launch_fast_0_0(b, io, st);
}

// generic kernel (or, for the humanoids, the fast pair): the kernels of each group width live in their own translation unit (cycle_tu.cuh)
int launch_cycle(iwbc_batch* b, const dev::IO& io, cudaStream_t st, int minb = 0)
{
    const bool use_fast = b->fast_ok && b->fast_enabled && b->G > 32 && !io.dbg_M && !io.dbg_cycles;
    if (use_fast)
        return launch_fast(b, io, st);
    if (b->G == 32)
        return minb == 4 ? iwbc_launch_cycle_32_4_4(b, io, st) : iwbc_launch_cycle_32_4_3(b, io, st);
    if (b->G == 64)
        return iwbc_launch_cycle_64(b, io, st);
    if (b->G == 96)
        return iwbc_launch_cycle_96(b, io, st);
    return iwbc_launch_cycle_128(b, io, st);
}

// choose the thread-group width for this problem shape
void choose_group(const iwbc_problem_t& P, int n, int& G, int& IPC)
{
    // (a 16-wide group with 8 instances per CTA was measured for the small arms -- Franka n = 9, Tiago n = 12: the kernel then needs
    // 218 registers, two CTAs per SM instead of three, and runs 30 % SLOWER: 26.8 M vs 38.4 M Franka updates/s.)
    if (n <= 32 && P.nj * 16 <= 8 * 32) {
        G = 32;
        IPC = 4;
        // (sub-warp groups -- 16 or 8 lanes per instance, 8 .. 24 instances per CTA -- were measured in round 2 with this generic code:
        // Franka 44.6 M/s at (16, 8) against 39.8 M/s, every other shape and every Tiago shape slower: profiles/r02_arm_shapes.md)
    }
    else if (n <= 64 && P.nj * 16 <= 8 * 64) {
        G = 64;
        IPC = 1;
    }
    else if (n <= 96 && P.nj * 16 <= 8 * 96) {
        G = 96;
        IPC = 1;
    }
    else {
        G = 128;
        IPC = 1;
    }
}

void stash_current(iwbc_batch* b)
{
    if (b->shape_mask == ~0u)
        return;
    iwbc_mask_state* slot = nullptr;
    for (auto& m : b->mask_cache)
        if (m.mask == b->shape_mask)
            slot = &m;
    if (!slot) {
        b->mask_cache.emplace_back();
        slot = &b->mask_cache.back();
    }
    slot->mask = b->shape_mask, slot->revision = b->shape_revision;
    slot->shape = b->shape, slot->fshape = b->fshape, slot->d_fplan = b->d_fplan;
    slot->G = b->G, slot->IPC = b->IPC, slot->grid = b->grid, slot->fast_grid = b->fast_grid, slot->slow_grid = b->slow_grid;
    slot->smem_bytes = b->smem_bytes, slot->fast_smem = b->fast_smem, slot->slow_smem = b->slow_smem, slot->fast_ok = b->fast_ok;
}

int prepare_shape(iwbc_batch* b, unsigned mask)
{
    const iwbc_problem_t& P = b->prob->cs.prob;
    if (b->shape_mask == mask && b->shape_revision == b->prob->revision)
        return IWBC_OK;
    // ---- another contact mask: park the current state, take the mask's own from the cache when it is still valid
    if (b->shape_mask != mask) {
        stash_current(b);
        b->d_fplan = nullptr;
        for (auto& m : b->mask_cache)
            if (m.mask == mask) {
                b->d_fplan = m.d_fplan; // (reused as the upload target even when the revision is stale)
                if (m.revision == b->prob->revision) {
                    b->shape = m.shape, b->fshape = m.fshape;
                    b->G = m.G, b->IPC = m.IPC, b->grid = m.grid, b->fast_grid = m.fast_grid, b->slow_grid = m.slow_grid;
                    b->smem_bytes = m.smem_bytes, b->fast_smem = m.fast_smem, b->slow_smem = m.slow_smem, b->fast_ok = m.fast_ok;
                    b->shape_mask = mask, b->shape_revision = m.revision;
                    b->plan_dirty = false;
                    return IWBC_OK;
                }
            }
    }
    dev::Shape& S = b->shape;
    std::memset(&S, 0, sizeof(S));
    host::QPShape qs = host::qp_shape(P, mask);
    S.n = qs.n;
    S.nk = qs.nk;
    S.me = qs.me;
    S.mi = qs.mi;
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
    S.rows_dense = 0;
    for (int k = 0; k < P.nl1; ++k)
        if (P.l1[k].kind == IWBC_B_TASK && P.tasks[P.l1[k].index].type != IWBC_T_POSTURE && P.tasks[P.l1[k].index].type != IWBC_T_TORQUE && P.tasks[P.l1[k].index].type != IWBC_T_COP)
            S.rows_dense += P.tasks[P.l1[k].index].nrows;
    S.maxdepth = 0;
    for (int j = 0; j < P.nj; ++j)
        S.maxdepth = std::max(S.maxdepth, P.depth[j]);
    S.ldJ = S.n | 1;
    S.ldM = P.nv | 1;
    dev::Sizes z;
    z.nv = P.nv, z.nq = P.nq, z.nj = P.nj, z.nf = P.nframes, z.nc = P.ncontacts, z.n = S.n, z.mi = S.mi;
    z.rows_dense = S.rows_dense, z.nref = P.nref, z.ldJ = S.ldJ, z.ldM = S.ldM;
    z.S = nullptr, z.P = nullptr;
    dev::make_layout(z, S);
    choose_group(P, S.n, b->G, b->IPC);
    b->smem_bytes = (size_t)S.total * sizeof(double) * b->IPC;
    b->grid = 0;
    b->shape_mask = mask;
    b->shape_revision = b->prob->revision;
    b->fast_ok = b->G > 32 && dev::build_fast_plan(P, S, 128, b->fshape, b->fplan);
    b->fast_smem = (size_t)b->fshape.total * sizeof(double);
    b->slow_smem = (size_t)b->fshape.s_total * sizeof(double);
    b->fast_grid = 0;
    b->slow_grid = 0;
    b->plan_dirty = true;
    return IWBC_OK;
}

// device copies of the problem (once per revision) and of the current mask's plan (once per (revision, mask)); stream-ordered
int upload_constants(iwbc_batch* b, cudaStream_t st)
{
    bool fresh = false;
    if (b->revision != b->prob->revision) {
        CUDA_TRY(cudaMemcpyAsync(b->d_prob, &b->prob->cs.prob, sizeof(iwbc_problem_t), cudaMemcpyHostToDevice, st));
        b->revision = b->prob->revision;
        b->dprob_revision = b->prob->revision;
        fresh = true;
    }
    if (b->plan_dirty) {
        if (!b->d_fplan)
            CUDA_TRY(cudaMalloc((void**)&b->d_fplan, sizeof(dev::FPlan)));
        // (pageable source: the copy is staged before the call returns; kernels of this mask launched earlier on this stream
        // still read the previous contents, later ones the new)
        CUDA_TRY(cudaMemcpyAsync(b->d_fplan, &b->fplan, sizeof(dev::FPlan), cudaMemcpyHostToDevice, st));
        b->plan_dirty = false;
        fresh = true;
    }
    if (fresh) // once per (revision, mask): later launches of this handle on OTHER streams must find the copies in place
        CUDA_TRY(cudaStreamSynchronize(st));
    return IWBC_OK;
}

} // namespace

// one pass of the cycle over io's instances (all of them, or io.list) with one contact mask; shared with behavior.cu
int iwbc_run_cycle(iwbc_batch* b, const dev::IO& io, unsigned mask, cudaStream_t st)
{
    if (b->prob->cs.prob.stabilizer_activated && !b->sensor_call) {
        // HumanoidPosTracker::update asserts its sensors when CONTROLLER.stabilizer.activated (humanoid_pos_tracker.cpp:180-184)
        set_error("the stabilizer needs the LF force: this controller has CONTROLLER.stabilizer.activated = true, call iwbc_batch_update_sensors"
                  " (or load it with stabilizer: {activated: false}, as the reference's open-loop tests do)");
        return IWBC_ERR_ARG;
    }
    if (b->prob->cs.torque_collision_detection) {
        // TalosPosTracker::update asserts "joints_torque" when CONTROLLER.collision_detection.activated (talos_pos_tracker.cpp:127-129);
        // the torque-discrepancy monitor is not part of the batched path
        set_error("torque collision detection requires torque sensor data: this controller has CONTROLLER.collision_detection.activated = true;"
                  " the monitor is outside the batched path -- load the controller with collision_detection: {activated: false}, as the reference's"
                  " controller-only tests do (tests/test_determinism.cpp:68)");
        return IWBC_ERR_UNSUPPORTED;
    }
    int rc = prepare_shape(b, mask);
    if (rc)
        return rc;
    rc = upload_constants(b, st);
    if (rc)
        return rc;
    if (b->G == 32 && !io.dbg_M && !io.dbg_cycles && b->shape.total <= 1792) {
        // Large arm batches run one THREAD per instance (k_cycle_tpi): a warp instruction advances 32 instances, lane efficiency
        // 27 / 32 instead of 9 / 32.  Measured (profiles/r02_tpi_*.txt): Franka 49.6 -> 141 M/s, Tiago 6.1 -> 40 M/s at 65 536 states,
        // four CTAs of 128 threads per SM (128 registers; 3 CTAs: 105 / 30 M/s, 5: 136 / 39, 6: 134 / 38).  A thread is ~30x slower
        // than a warp on one instance, so small batches (latency, debugging) stay on the warp-per-instance kernel below (measured
        // crossover, profiles/r02_tpi_crossover.txt: Franka ~9 k states, Tiago ~11 k).
        static const int tpi_min = [] { const char* e = std::getenv("IWBC_ARM_TPI_MIN_BATCH"); return e ? std::atoi(e) : 12288; }();
        if (io.B >= tpi_min) {
            dev::IO iop = io;
            if (io.in_list)
                iop.list = io.in_list, iop.list_count = io.in_count;
            static const int tpi_ctas = [] { const char* e = std::getenv("IWBC_TPI_CTAS"); return e ? std::atoi(e) : 4; }(); // CTAs per SM (tuning aid)
            const int grid = std::min(b->sm_count * tpi_ctas, (io.B + 127) / 128);
            return b->shape.total <= 1024 ? iwbc_launch_cycle_tpi_1024(b, iop, grid, st) : iwbc_launch_cycle_tpi_1792(b, iop, grid, st);
        }
    }
    if (b->G == 32) {
        // The warp-per-instance kernel is latency bound (dependent shared-memory load -> FP64 chains): resident warps are what it
        // needs.  Four CTAs per SM (<= 128 registers, 16 warps) when the instance's shared memory allows it (Franka: 36 kB per CTA),
        // three otherwise (Tiago: 60 kB).  Measured: Franka 29.0 -> 49.9 M/s from 3 to 4 CTAs, 38.3 at 5 (96 registers: spills).
        static const int minb = [] { const char* e = std::getenv("IWBC_ARM_MINB"); return e ? std::atoi(e) : 0; }(); // tuning aid
        // (five / six CTAs per SM -- 96 / 80 registers -- change nothing any more once the phases are inlined: Franka 50.8 / 48.7 M/s
        // against 49.6, Tiago 6.13 / 6.14 against 6.14: profiles/README.md)
        const int fit = (int)(233472 / (b->smem_bytes + 1024));
        const int want = minb ? minb : (fit >= 4 ? 4 : 3);
        return launch_cycle(b, io, st, want >= 4 && fit >= 4 ? 4 : 3);
    }
    return launch_cycle(b, io, st);
}

namespace {
int run_cycle(iwbc_batch* b, const dev::IO& io, unsigned mask, cudaStream_t st) { return iwbc_run_cycle(b, io, mask, st); }

template <class T>
int dalloc(T** p, size_t count)
{
    CUDA_TRY(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    return IWBC_OK;
}

} // namespace

int iwbc_raise_smem_limit(const void* func, int device, size_t bytes, size_t* max_set)
{
    std::lock_guard<std::mutex> lock(g_attr_mutex);
    size_t& cur = max_set[device & 15];
    if (bytes > cur) {
        CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
    return IWBC_OK;
}

extern "C" {

iwbc_batch* iwbc_batch_create(const iwbc_problem* p, int device, int max_batch)
{
    if (!p || max_batch <= 0) {
        set_error("iwbc_batch_create: bad argument");
        return nullptr;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || device < 0 || device >= ndev) {
        set_error(std::string("no usable CUDA device ") + std::to_string(device) + " (" + (e != cudaSuccess ? cudaGetErrorString(e) : "out of range")
            + "): the batched cycle has no CPU fallback");
        return nullptr;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) {
        set_error("device " + std::to_string(device) + " is sm_" + std::to_string(prop.major * 10 + prop.minor) + "; this library is built for sm_100a only");
        return nullptr;
    }
    cudaSetDevice(device);
    auto* b = new iwbc_batch;
    b->prob = p;
    b->device = device;
    b->max_batch = max_batch;
    b->sm_count = prop.multiProcessorCount;
    const iwbc_problem_t& P = p->cs.prob;
    const size_t B = (size_t)max_batch;
    bool ok = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&b->ev0) == cudaSuccess && cudaEventCreate(&b->ev1) == cudaSuccess;
    ok = ok && dalloc(&b->d_slow_list, B) == 0 && dalloc(&b->d_slow_count, 16) == 0 && dalloc(&b->d_slow_code, 65536) == 0;
    ok = ok && cudaStreamCreateWithFlags(&b->aux[0], cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&b->aux[1], cudaStreamNonBlocking) == cudaSuccess
        && cudaEventCreateWithFlags(&b->ev_const, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&b->s_in, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&b->s_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; k < 16 && ok; ++k)
        ok = cudaEventCreateWithFlags(&b->ev_in[k], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&b->ev_done[k], cudaEventDisableTiming) == cudaSuccess;
    {
        const char* e = getenv("IWBC_NO_FAST");
        b->fast_enabled = !(e && e[0] == '1');
    }
    ok = ok && dalloc(&b->d_prof, 48) == 0 && dalloc(&b->d_prob, 1) == 0 && dalloc(&b->d_q, B * P.nq) == 0 && dalloc(&b->d_v, B * P.nv) == 0
        && dalloc(&b->d_refs, B * P.nref) == 0 && dalloc(&b->d_tau, B * P.na) == 0 && dalloc(&b->d_dv, B * P.nv) == 0
        && dalloc(&b->d_f, B * 12 * P.ncontacts) == 0 && dalloc(&b->d_qn, B * P.nq) == 0 && dalloc(&b->d_vn, B * P.nv) == 0
        && dalloc(&b->d_qc, B * P.nv) == 0 && dalloc(&b->d_mom, B * 3) == 0 && dalloc(&b->d_status, B) == 0 && dalloc(&b->d_iters, B) == 0
        && dalloc(&b->d_active, B * IWBC_ACTIVE_WORDS) == 0;
    if (!ok) {
        if (std::string(iwbc_last_error()).empty())
            set_error("iwbc_batch_create: CUDA allocation failed");
        iwbc_batch_free(b);
        return nullptr;
    }
    return b;
}

void iwbc_batch_free(iwbc_batch* b)
{
    if (!b)
        return;
    cudaSetDevice(b->device);
    if (b->stream)
        cudaStreamSynchronize(b->stream);
    for (auto& m : b->mask_cache)
        if (m.d_fplan && m.d_fplan != b->d_fplan)
            cudaFree(m.d_fplan);
    void* ptrs[] = {b->d_sensors, b->d_phase, b->d_cmask, b->d_scratch, b->d_slow_code, b->d_fplan, b->d_slow_list, b->d_slow_count, b->d_prof, b->d_prob, b->d_q, b->d_v, b->d_refs, b->d_tau, b->d_dv, b->d_f, b->d_qn, b->d_vn, b->d_qc, b->d_mom, b->d_status, b->d_iters, b->d_active};
    for (void* p : ptrs)
        if (p)
            cudaFree(p);
    for (int k = 0; k < 2; ++k)
        if (b->aux[k]) {
            cudaStreamSynchronize(b->aux[k]);
            cudaStreamDestroy(b->aux[k]);
        }
    if (b->ev_const)
        cudaEventDestroy(b->ev_const);
    for (cudaStream_t q : {b->s_in, b->s_out})
        if (q) {
            cudaStreamSynchronize(q);
            cudaStreamDestroy(q);
        }
    for (int k = 0; k < 16; ++k) {
        if (b->ev_in[k])
            cudaEventDestroy(b->ev_in[k]);
        if (b->ev_done[k])
            cudaEventDestroy(b->ev_done[k]);
    }
    if (b->ev0)
        cudaEventDestroy(b->ev0);
    if (b->ev1)
        cudaEventDestroy(b->ev1);
    if (b->stream)
        cudaStreamDestroy(b->stream);
    delete b;
}

int iwbc_batch_update_device(iwbc_batch* b, int B, const double* q, const double* v, const double* refs, long refs_stride, unsigned contact_mask,
    const iwbc_results_t* out, void* stream)
{
    if (!b || !q || !v || !refs || !out || B < 0) {
        set_error("iwbc_batch_update_device: bad argument");
        return IWBC_ERR_ARG;
    }
    if (B > b->max_batch) { // the continuation list, the hand-over scratch and the counters are sized for max_batch
        set_error("iwbc_batch_update_device: B = " + std::to_string(B) + " exceeds max_batch = " + std::to_string(b->max_batch));
        return IWBC_ERR_ARG;
    }
    if (B == 0)
        return IWBC_OK;
    CUDA_TRY(cudaSetDevice(b->device));
    dev::IO io;
    std::memset(&io, 0, sizeof(io));
    io.B = B;
    io.q = q, io.v = v, io.refs = refs, io.refs_stride = refs_stride;
    io.tau = out->tau, io.dv = out->dv, io.f = out->f, io.q_next = out->q_next, io.v_next = out->v_next, io.q_cmd = out->q_cmd;
    io.momentum = out->momentum, io.status = out->status, io.iters = out->iters, io.active = out->active;
    return run_cycle(b, io, contact_mask, stream ? (cudaStream_t)stream : b->stream);
}

int iwbc_batch_rollout_device(iwbc_batch* b, int B, double* q, double* v, const double* refs, long refs_stride, unsigned contact_mask, int ticks,
    const iwbc_results_t* out_last, void* stream)
{
    if (!b || !q || !v || !refs || B < 0 || ticks < 0) {
        set_error("iwbc_batch_rollout_device: bad argument");
        return IWBC_ERR_ARG;
    }
    if (B > b->max_batch) {
        set_error("iwbc_batch_rollout_device: B = " + std::to_string(B) + " exceeds max_batch = " + std::to_string(b->max_batch));
        return IWBC_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(b->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    for (int t = 0; t < ticks; ++t) {
        dev::IO io;
        std::memset(&io, 0, sizeof(io));
        io.B = B;
        io.q = q, io.v = v, io.refs = refs, io.refs_stride = refs_stride;
        io.q_rw = q, io.v_rw = v;
        // A tick whose QP is not solved to optimality does not advance that instance's state (the reference throws before
        // q_tsid_ / v_tsid_ are touched, controller.cpp:249-256) and its status is kept for the rest of the call (sticky).
        io.sticky_status = t > 0;
        if (out_last) {
            io.status = out_last->status, io.iters = out_last->iters;
            if (t == ticks - 1) {
                io.tau = out_last->tau, io.dv = out_last->dv, io.f = out_last->f, io.q_cmd = out_last->q_cmd;
                io.momentum = out_last->momentum, io.active = out_last->active;
            }
        }
        int rc = run_cycle(b, io, contact_mask, st);
        if (rc)
            return rc;
    }
    return IWBC_OK;
}

int iwbc_batch_update(iwbc_batch* b, int B, const double* q, const double* v, const double* refs, long refs_stride, const unsigned* contact_mask,
    const iwbc_results_t* out)
{
    if (!b || !q || !v || !refs || !out || B < 0) {
        set_error("iwbc_batch_update: bad argument");
        return IWBC_ERR_ARG;
    }
    if (B > b->max_batch) {
        set_error("iwbc_batch_update: B = " + std::to_string(B) + " exceeds max_batch = " + std::to_string(b->max_batch));
        return IWBC_ERR_ARG;
    }
    if (refs_stride != 0 && refs_stride != b->prob->cs.prob.nref) {
        set_error("iwbc_batch_update: refs_stride must be 0 (shared) or nref");
        return IWBC_ERR_ARG;
    }
    if (B == 0)
        return IWBC_OK;
    CUDA_TRY(cudaSetDevice(b->device));
    const iwbc_problem_t& P = b->prob->cs.prob;
    cudaStream_t st = b->stream;
    const unsigned dflt = iwbc_problem_default_contact_mask(b->prob);
    // ---- bucket by contact mask (every launch is shape-uniform): instance order is restored on the way out
    bool uniform = true;
    if (contact_mask)
        for (int i = 1; i < B; ++i)
            if (contact_mask[i] != contact_mask[0]) {
                uniform = false;
                break;
            }
    const size_t nq = P.nq, nv = P.nv, na = P.na, nf = 12 * P.ncontacts, nref = P.nref;
    auto copy_out = [&](int first, int count, const int* perm, cudaStream_t st) -> int {
        // device slot s in [first, first+count) holds instance perm[s] (or s when perm == nullptr)
        auto back = [&](void* host, const void* devp, size_t elem_bytes, size_t per) -> int {
            if (!host)
                return IWBC_OK;
            if (!perm) {
                CUDA_TRY(cudaMemcpyAsync((char*)host + first * per * elem_bytes, (const char*)devp + first * per * elem_bytes, count * per * elem_bytes,
                    cudaMemcpyDeviceToHost, st));
            }
            else {
                for (int s = first; s < first + count; ++s)
                    CUDA_TRY(cudaMemcpyAsync((char*)host + (size_t)perm[s] * per * elem_bytes, (const char*)devp + (size_t)s * per * elem_bytes,
                        per * elem_bytes, cudaMemcpyDeviceToHost, st));
            }
            return IWBC_OK;
        };
        int rc = 0;
        rc |= back(out->tau, b->d_tau, 8, na);
        rc |= back(out->dv, b->d_dv, 8, nv);
        rc |= back(out->f, b->d_f, 8, nf);
        rc |= back(out->q_next, b->d_qn, 8, nq);
        rc |= back(out->v_next, b->d_vn, 8, nv);
        rc |= back(out->q_cmd, b->d_qc, 8, nv);
        rc |= back(out->momentum, b->d_mom, 8, 3);
        rc |= back(out->status, b->d_status, 4, 1);
        rc |= back(out->iters, b->d_iters, 4, 1);
        rc |= back(out->active, b->d_active, 4, IWBC_ACTIVE_WORDS);
        return rc ? IWBC_ERR_CUDA : IWBC_OK;
    };
    auto dev_results = [&](size_t first) {
        iwbc_results_t r;
        r.tau = out->tau ? b->d_tau + first * na : nullptr;
        r.dv = out->dv ? b->d_dv + first * nv : nullptr;
        r.f = out->f ? b->d_f + first * nf : nullptr;
        r.q_next = out->q_next ? b->d_qn + first * nq : nullptr;
        r.v_next = out->v_next ? b->d_vn + first * nv : nullptr;
        r.q_cmd = out->q_cmd ? b->d_qc + first * nv : nullptr;
        r.momentum = out->momentum ? b->d_mom + first * 3 : nullptr;
        r.status = out->status ? b->d_status + first : nullptr;
        r.iters = out->iters ? b->d_iters + first : nullptr;
        r.active = out->active ? b->d_active + first * IWBC_ACTIVE_WORDS : nullptr;
        return r;
    };
    // contiguous runs of equal contact mask (one run when the mask is uniform; the caller may have bucketed a mixed batch)
    struct Run {
        int first, count;
        unsigned mask;
    };
    std::vector<Run> runs;
    bool grouped = true;
    if (uniform)
        runs.push_back({0, B, contact_mask ? contact_mask[0] : dflt});
    else {
        for (int i = 0; i < B && grouped; ++i)
            if (i == 0 || contact_mask[i] != contact_mask[i - 1]) {
                for (const Run& r : runs)
                    if (r.mask == contact_mask[i])
                        grouped = false;
                runs.push_back({i, 0, contact_mask[i]});
            }
        if (grouped) {
            for (size_t k = 0; k < runs.size(); ++k)
                runs[k].count = (k + 1 < runs.size() ? runs[k + 1].first : B) - runs[k].first;
        }
    }
    if (grouped && B >= 8192 && runs.size() <= 4 && !b->timing && !b->profiling) {
        // Pipeline (needs pinned host buffers to be effective).  Uploads queue back to back on their own stream (the copy engine
        // never waits for a kernel or a download); the kernels of successive chunks alternate between two compute streams (their
        // tails overlap); downloads follow on a third stream; events chain the three per chunk.  Every contact mask has its own
        // shape parameters and device plan, so runs of different masks overlap like chunks of one run.
        int c = 0; // global chunk index (<= 16: ev_in / ev_done)
        const int cap = std::min(b->max_batch, 65536);
        for (size_t ri = 0; ri < runs.size(); ++ri) {
            const Run& run = runs[ri];
            int rc = prepare_shape(b, run.mask);
            if (rc)
                return rc;
            rc = upload_constants(b, st);
            if (rc)
                return rc;
            CUDA_TRY(cudaEventRecord(b->ev_const, st));
            CUDA_TRY(cudaStreamWaitEvent(b->aux[0], b->ev_const, 0));
            CUDA_TRY(cudaStreamWaitEvent(b->aux[1], b->ev_const, 0));
            if (ri == 0)
                CUDA_TRY(cudaStreamWaitEvent(b->s_in, b->ev_const, 0)); // (the previous call's kernels are done: the call synchronises before returning)
            // Chunk plan: sizes are multiples of the persistent grid's instance capacity (no partially filled last round), a short
            // first chunk (its upload is the pipeline's fill: nothing overlaps it) and whatever is left at the end.
            // (thread-per-instance arms: a chunk needs tens of thousands of instances to fill the GPU -- 32 instances per warp)
            const bool tpi_arm = b->G == 32 && b->shape.total <= 1792;
            const int unit = tpi_arm ? 32 * b->sm_count : b->fast_ok && b->fast_grid > 0 ? b->fast_grid : (b->grid > 0 ? b->grid * b->IPC : 4 * b->sm_count);
            // (IWBC_PIPE="first,mid" in units of the grid capacity: tuning aid.  Humanoids at 4 CTAs / SM -- unit 592 -- measured on
            // Talos double support, 16 384 states: 4,10 -> 4.32 M/s, 2,10 -> 4.14, 3,8 -> 4.31, 1,6 -> 4.40, 2,4 -> 4.42, 2,6 -> 4.52)
            // Thread-per-instance arms -- unit 4 736 -- at 65 536 states (profiles/r02d_pipe_arms.txt): Franka 4,10 -> 48.0 M/s, 3,4 -> 59.1,
            // 3,6 -> 54.7; Tiago 4,10 -> 23.3, 3,4 -> 23.0, 3,6 -> 23.8 (its kernel needs larger chunks to fill the GPU).
            int u_first = tpi_arm ? 3 : 2, u_mid = tpi_arm ? (b->shape.total <= 1024 ? 4 : 6) : 6;
            if (const char* e = std::getenv("IWBC_PIPE"))
                std::sscanf(e, "%d,%d", &u_first, &u_mid);
            const int budget = std::max(1, (15 - c) / (int)(runs.size() - ri)); // chunks this run may use
            std::vector<int> bounds;
            bounds.push_back(0);
            bounds.push_back(std::min(run.count, std::max(1, ri == 0 ? u_first : u_mid) * unit));
            int step = std::max(1, u_mid) * unit;
            if ((run.count - bounds.back() + step - 1) / step > budget - 1) // long run: fewer, larger chunks (still multiples of the unit)
                step = ((run.count - bounds.back() + (budget - 1) * unit - 1) / ((budget - 1 > 0 ? budget - 1 : 1) * unit)) * unit;
            while (bounds.back() < run.count)
                bounds.push_back(std::min(run.count, bounds.back() + std::max(step, unit)));
            for (size_t k = 0; k + 1 < bounds.size(); ++k, ++c) {
                if (c > 14) {
                    set_error("iwbc_batch_update: internal error (pipeline chunk plan)");
                    return IWBC_ERR_ARG;
                }
                const size_t first = (size_t)run.first + bounds[k], cnt = bounds[k + 1] - bounds[k];
                cudaStream_t cs = b->aux[c & 1];
                CUDA_TRY(cudaMemcpyAsync(b->d_q + first * nq, q + first * nq, cnt * nq * 8, cudaMemcpyHostToDevice, b->s_in));
                CUDA_TRY(cudaMemcpyAsync(b->d_v + first * nv, v + first * nv, cnt * nv * 8, cudaMemcpyHostToDevice, b->s_in));
                if (refs_stride)
                    CUDA_TRY(cudaMemcpyAsync(b->d_refs + first * nref, refs + first * nref, cnt * nref * 8, cudaMemcpyHostToDevice, b->s_in));
                else if (c == 0) // shared reference vector
                    CUDA_TRY(cudaMemcpyAsync(b->d_refs, refs, nref * 8, cudaMemcpyHostToDevice, b->s_in));
                CUDA_TRY(cudaEventRecord(b->ev_in[c], b->s_in));
                CUDA_TRY(cudaStreamWaitEvent(cs, b->ev_in[c], 0));
                iwbc_results_t r = dev_results(first);
                b->slot = c & 7; // chunks c and c + 8 share their counters: same stream, ordered
                b->list_off = (int)first;
                if (B <= cap) { // hand-over images: the chunk's own range of the scratch
                    b->scratch_off = (int)first;
                    b->scratch_len = (int)cnt;
                }
                else { // one half per compute stream: chunks on the same stream are ordered
                    b->scratch_len = cap / 2;
                    b->scratch_off = (c & 1) * b->scratch_len;
                }
                rc = iwbc_batch_update_device(b, (int)cnt, b->d_q + first * nq, b->d_v + first * nv, b->d_refs + (refs_stride ? first * nref : 0),
                    refs_stride, run.mask, &r, cs);
                b->slot = 0;
                b->list_off = 0;
                b->scratch_len = 0;
                if (rc)
                    return rc;
                CUDA_TRY(cudaEventRecord(b->ev_done[c], cs));
                CUDA_TRY(cudaStreamWaitEvent(b->s_out, b->ev_done[c], 0));
                rc = copy_out((int)first, (int)cnt, nullptr, b->s_out);
                if (rc)
                    return rc;
            }
        }
        CUDA_TRY(cudaStreamSynchronize(b->s_out)); // (every upload and kernel precedes the last download)
        CUDA_TRY(cudaStreamSynchronize(b->aux[0]));
        CUDA_TRY(cudaStreamSynchronize(b->aux[1]));
        return IWBC_OK;
    }
    if (uniform) {
        CUDA_TRY(cudaMemcpyAsync(b->d_q, q, B * nq * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(b->d_v, v, B * nv * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(b->d_refs, refs, (refs_stride ? (size_t)B : 1) * nref * 8, cudaMemcpyHostToDevice, st));
        iwbc_results_t r = dev_results(0);
        int rc = iwbc_batch_update_device(b, B, b->d_q, b->d_v, b->d_refs, refs_stride, contact_mask ? contact_mask[0] : dflt, &r, st);
        if (rc)
            return rc;
        rc = copy_out(0, B, nullptr, st);
        if (rc)
            return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
        return IWBC_OK;
    }
    // mixed masks.  If equal masks already form contiguous runs (the caller bucketed the batch) the runs are launched in
    // place; otherwise: stable sort of instance indices by mask, gather on the host, one launch per mask, results copied
    // back in bulk and scattered on the host.
    const double *sq = q, *sv = v, *sr = refs;
    b->perm.resize(B);
    for (int i = 0; i < B; ++i)
        b->perm[i] = i;
    if (!grouped) {
        std::stable_sort(b->perm.begin(), b->perm.end(), [&](int a, int c) { return contact_mask[a] < contact_mask[c]; });
        b->hq.resize(B * nq);
        b->hv.resize(B * nv);
        if (refs_stride)
            b->hr.resize(B * nref);
        for (int s = 0; s < B; ++s) {
            const int i = b->perm[s];
            std::memcpy(&b->hq[s * nq], q + i * nq, nq * 8);
            std::memcpy(&b->hv[s * nv], v + i * nv, nv * 8);
            if (refs_stride)
                std::memcpy(&b->hr[s * nref], refs + (size_t)i * nref, nref * 8);
        }
        sq = b->hq.data(), sv = b->hv.data();
        if (refs_stride)
            sr = b->hr.data();
    }
    CUDA_TRY(cudaMemcpyAsync(b->d_q, sq, B * nq * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_v, sv, B * nv * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_refs, sr, (refs_stride ? (size_t)B : 1) * nref * 8, cudaMemcpyHostToDevice, st));
    int first = 0;
    while (first < B) {
        const unsigned m = contact_mask[b->perm[first]];
        int last = first;
        while (last < B && contact_mask[b->perm[last]] == m)
            ++last;
        iwbc_results_t r = dev_results(first);
        int rc = iwbc_batch_update_device(b, last - first, b->d_q + first * nq, b->d_v + first * nv, b->d_refs + (refs_stride ? first * nref : 0),
            refs_stride, m, &r, st);
        if (rc)
            return rc;
        first = last;
    }
    if (grouped) {
        int rc = copy_out(0, B, nullptr, st);
        if (rc)
            return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
        return IWBC_OK;
    }
    // bulk D2H into staging, then scatter slot s -> instance perm[s]
    struct Field {
        void* host;
        const void* dev;
        size_t bytes;
    };
    const Field fields[] = {{out->tau, b->d_tau, na * 8}, {out->dv, b->d_dv, nv * 8}, {out->f, b->d_f, nf * 8}, {out->q_next, b->d_qn, nq * 8},
        {out->v_next, b->d_vn, nv * 8}, {out->q_cmd, b->d_qc, nv * 8}, {out->momentum, b->d_mom, 3 * 8}, {out->status, b->d_status, 4},
        {out->iters, b->d_iters, 4}, {out->active, b->d_active, IWBC_ACTIVE_WORDS * 4}};
    size_t total = 0;
    for (const Field& f : fields)
        if (f.host)
            total += f.bytes * B;
    b->stage.resize(total);
    size_t off = 0;
    for (const Field& f : fields)
        if (f.host && f.bytes) {
            CUDA_TRY(cudaMemcpyAsync(b->stage.data() + off, f.dev, f.bytes * B, cudaMemcpyDeviceToHost, st));
            off += f.bytes * B;
        }
    CUDA_TRY(cudaStreamSynchronize(st));
    off = 0;
    for (const Field& f : fields)
        if (f.host && f.bytes) {
            for (int s2 = 0; s2 < B; ++s2)
                std::memcpy((char*)f.host + (size_t)b->perm[s2] * f.bytes, b->stage.data() + off + (size_t)s2 * f.bytes, f.bytes);
            off += f.bytes * B;
        }
    return IWBC_OK;
}

int iwbc_batch_debug_cycles(const iwbc_batch* b, long long* out7)
{
    if (!b || !out7)
        return IWBC_ERR_ARG;
    for (int k = 0; k < 7; ++k)
        out7[k] = b->last_cycles[k];
    return IWBC_OK;
}

int iwbc_batch_phase_profile(iwbc_batch* b, int enable, unsigned long long* out7)
{
    if (!b)
        return IWBC_ERR_ARG;
    CUDA_TRY(cudaSetDevice(b->device));
    CUDA_TRY(cudaDeviceSynchronize());
    if (out7)
        CUDA_TRY(cudaMemcpy(out7, b->d_prof, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(b->loop_prof, b->d_prof + 16, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemset(b->d_prof, 0, 48 * sizeof(unsigned long long)));
    b->profiling = enable != 0;
    return IWBC_OK;
}

int iwbc_batch_loop_profile(const iwbc_batch* b, unsigned long long* out8)
{
    if (!b || !out8)
        return IWBC_ERR_ARG;
    for (int k = 0; k < 16; ++k)
        out8[k] = b->loop_prof[k];
    return IWBC_OK;
}

int iwbc_batch_launch_info(const iwbc_batch* b, int* G, int* ipc, int* grid, long* smem_bytes)
{
    if (!b)
        return IWBC_ERR_ARG;
    if (G)
        *G = b->G;
    if (ipc)
        *ipc = b->IPC;
    if (grid)
        *grid = b->grid;
    if (smem_bytes)
        *smem_bytes = (long)b->smem_bytes;
    return IWBC_OK;
}

int iwbc_batch_fast_info(const iwbc_batch* b, long* out6)
{
    if (!b || !out6)
        return IWBC_ERR_ARG;
    out6[0] = b->fast_ok && b->fast_enabled;
    out6[1] = b->fast_grid;
    out6[2] = (long)b->fast_smem;
    out6[3] = b->slow_grid;
    out6[4] = (long)b->slow_smem;
    out6[5] = IWBC_FAST_THREADS;
    return IWBC_OK;
}

double iwbc_measure_fp64_tflops(int device)
{
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("iwbc_measure_fp64_tflops: no usable CUDA device");
        return -1.0;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    double* d = nullptr;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 1 << 14, blocks = prop.multiProcessorCount * 8, threads = 256;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dev::k_fp64_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best)
            best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}

long iwbc_batch_launch_count(const iwbc_batch* b) { return b ? b->launches : 0; }

int iwbc_batch_enable_timing(iwbc_batch* b, int on)
{
    if (!b)
        return IWBC_ERR_ARG;
    b->timing = on != 0;
    return IWBC_OK;
}

double iwbc_batch_kernel_ms(iwbc_batch* b, int reset)
{
    if (!b)
        return 0;
    double r = b->ms_n ? b->ms_sum / b->ms_n : 0.0;
    if (reset) {
        b->ms_sum = 0;
        b->ms_n = 0;
    }
    return r;
}

// stage dump of one instance through the generic kernel (any output may be NULL); A / bvec: the dense level-1 rows and right-hand sides
static int debug_dump(iwbc_batch* b, const double* q, const double* v, const double* ref, unsigned contact_mask, double* M, double* nle,
    double* H, double* g, double* x, double* A, double* bvec)
{
    if (!b || !q || !v || !ref) {
        set_error("iwbc_batch_debug_instance: bad argument");
        return IWBC_ERR_ARG;
    }
    CUDA_TRY(cudaSetDevice(b->device));
    const iwbc_problem_t& P = b->prob->cs.prob;
    host::QPShape qs = host::qp_shape(P, contact_mask);
    const size_t nv = P.nv, n = qs.n;
    size_t rows = 0;
    for (int k = 0; k < P.nl1; ++k)
        if (P.l1[k].kind == IWBC_B_TASK && P.tasks[P.l1[k].index].type != IWBC_T_POSTURE && P.tasks[P.l1[k].index].type != IWBC_T_TORQUE && P.tasks[P.l1[k].index].type != IWBC_T_COP)
            rows += P.tasks[P.l1[k].index].nrows;
    double *dM, *dn, *dH, *dg, *dx, *dA = nullptr, *db = nullptr;
    CUDA_TRY(cudaMalloc(&dM, nv * nv * 8));
    CUDA_TRY(cudaMalloc(&dn, nv * 8));
    CUDA_TRY(cudaMalloc(&dH, n * n * 8));
    CUDA_TRY(cudaMalloc(&dg, n * 8));
    CUDA_TRY(cudaMalloc(&dx, n * 8));
    if (A) {
        CUDA_TRY(cudaMalloc(&dA, std::max<size_t>(rows * nv, 1) * 8));
        CUDA_TRY(cudaMalloc(&db, std::max<size_t>(rows, 1) * 8));
    }
    cudaStream_t st = b->stream;
    CUDA_TRY(cudaMemcpyAsync(b->d_q, q, P.nq * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_v, v, nv * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_refs, ref, P.nref * 8, cudaMemcpyHostToDevice, st));
    dev::IO io;
    std::memset(&io, 0, sizeof(io));
    io.B = 1;
    io.q = b->d_q, io.v = b->d_v, io.refs = b->d_refs, io.refs_stride = 0;
    io.dbg_M = dM, io.dbg_nle = dn, io.dbg_H = dH, io.dbg_g = dg, io.dbg_x = dx, io.dbg_A = dA, io.dbg_b = db;
    long long* dcyc = nullptr;
    CUDA_TRY(cudaMalloc(&dcyc, 8 * sizeof(long long)));
    CUDA_TRY(cudaMemsetAsync(dcyc, 0, 8 * sizeof(long long), st));
    io.dbg_cycles = dcyc;
    int rc = run_cycle(b, io, contact_mask, st);
    if (rc == 0) {
        if (M)
            cudaMemcpyAsync(M, dM, nv * nv * 8, cudaMemcpyDeviceToHost, st);
        if (nle)
            cudaMemcpyAsync(nle, dn, nv * 8, cudaMemcpyDeviceToHost, st);
        if (H)
            cudaMemcpyAsync(H, dH, n * n * 8, cudaMemcpyDeviceToHost, st);
        if (g)
            cudaMemcpyAsync(g, dg, n * 8, cudaMemcpyDeviceToHost, st);
        if (x)
            cudaMemcpyAsync(x, dx, n * 8, cudaMemcpyDeviceToHost, st);
        if (A && rows) {
            cudaMemcpyAsync(A, dA, rows * nv * 8, cudaMemcpyDeviceToHost, st);
            cudaMemcpyAsync(bvec, db, rows * 8, cudaMemcpyDeviceToHost, st);
        }
        cudaMemcpyAsync(b->last_cycles, dcyc, 8 * sizeof(long long), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_error(std::string("CUDA error: ") + cudaGetErrorString(e));
            rc = IWBC_ERR_CUDA;
        }
    }
    cudaFree(dM), cudaFree(dn), cudaFree(dH), cudaFree(dg), cudaFree(dx), cudaFree(dcyc), cudaFree(dA), cudaFree(db);
    return rc;
}

int iwbc_batch_debug_instance(iwbc_batch* b, const double* q, const double* v, const double* ref, unsigned contact_mask, double* M, double* nle,
    double* H, double* g, double* x)
{
    return debug_dump(b, q, v, ref, contact_mask, M, nle, H, g, x, nullptr, nullptr);
}

int iwbc_batch_objective_value(iwbc_batch* b, const double* q, const double* v, const double* ref, unsigned contact_mask, double* objective)
{
    if (!b || !objective) {
        set_error("iwbc_batch_objective_value: bad argument");
        return IWBC_ERR_ARG;
    }
    const size_t n = host::qp_shape(b->prob->cs.prob, contact_mask).n;
    std::vector<double> H(n * n), g(n), x(n);
    int rc = debug_dump(b, q, v, ref, contact_mask, nullptr, nullptr, H.data(), g.data(), x.data(), nullptr, nullptr);
    if (rc)
        return rc;
    double f = 0; // eiquadprog's f_value at the solution: x'Hx/2 + g'x
    for (size_t i = 0; i < n; ++i) {
        double hx = 0;
        for (size_t j = 0; j < n; ++j)
            hx += H[i * n + j] * x[j];
        f += x[i] * (0.5 * hx + g[i]);
    }
    *objective = f;
    return IWBC_OK;
}

int iwbc_batch_task_cost(iwbc_batch* b, const double* q, const double* v, const double* ref, unsigned contact_mask, const char* task,
    const double* ddq, const double* tau, double* cost)
{
    if (!b || !q || !v || !ref || !task || !ddq || !cost) {
        set_error("iwbc_batch_task_cost: bad argument");
        return IWBC_ERR_ARG;
    }
    const iwbc_problem_t& P = b->prob->cs.prob;
    int t = -1;
    for (size_t i = 0; i < b->prob->cs.task_names.size(); ++i)
        if (b->prob->cs.task_names[i] == task)
            t = (int)i;
    if (t < 0 || P.tasks[t].level != 1) {
        set_error(std::string("iwbc_batch_task_cost: '") + task + "' is not a level-1 task of this controller");
        return IWBC_ERR_ARG;
    }
    const iwbc_task_t& T = P.tasks[t];
    const int nv = P.nv, na = P.na, nu = nv - na;
    double s2 = 0;
    if (T.type == IWBC_T_POSTURE) { // TaskJointPosture: S (ddq_a - a_des)
        const double* r = ref + T.ref_off;
        for (int i = 0; i < na; ++i)
            if (T.mask[i]) {
                const double a_des = -T.kp[0] * (q[P.nq - na + i] - r[i]) - T.kd[0] * (v[nv - na + i] - r[na + i]) + r[2 * na + i];
                s2 += (ddq[nu + i] - a_des) * (ddq[nu + i] - a_des);
            }
    }
    else if (T.type == IWBC_T_TORQUE) { // TaskActuationEquality: diag(w) S (tau - tau_ref)
        if (!tau) {
            set_error("iwbc_batch_task_cost: a torque task needs the joint torques of that solve");
            return IWBC_ERR_ARG;
        }
        const double* r = ref + T.ref_off;
        for (int i = 0; i < na; ++i)
            if (T.mask[i])
                s2 += r[na + i] * (tau[i] - r[i]) * r[na + i] * (tau[i] - r[i]);
    }
    else if (T.type == IWBC_T_COP) {
        set_error("iwbc_batch_task_cost: the cop task is a cost on the contact forces (the reference's cost() multiplies task matrices by ddq: not defined for it)");
        return IWBC_ERR_UNSUPPORTED;
    }
    else {
        if (!(T.weight > 0)) {
            set_error(std::string("iwbc_batch_task_cost: task '") + task + "' has weight 0: its rows are not formed");
            return IWBC_ERR_UNSUPPORTED;
        }
        int row0 = 0, rows = 0;
        for (int k = 0; k < P.nl1; ++k) {
            if (P.l1[k].kind != IWBC_B_TASK)
                continue;
            const iwbc_task_t& Tk = P.tasks[P.l1[k].index];
            if (Tk.type == IWBC_T_POSTURE || Tk.type == IWBC_T_TORQUE || Tk.type == IWBC_T_COP)
                continue;
            if (P.l1[k].index == t)
                row0 = rows;
            rows += Tk.nrows;
        }
        std::vector<double> A((size_t)std::max(rows, 1) * nv), bv(std::max(rows, 1));
        int rc = debug_dump(b, q, v, ref, contact_mask, nullptr, nullptr, nullptr, nullptr, nullptr, A.data(), bv.data());
        if (rc)
            return rc;
        const double sw = std::sqrt(T.weight); // (the rows are stored scaled by sqrt(weight); the reference's constraint is not)
        for (int r = row0; r < row0 + T.nrows; ++r) {
            double a = -bv[r];
            for (int k = 0; k < nv; ++k)
                a += A[(size_t)r * nv + k] * ddq[k];
            s2 += (a / sw) * (a / sw);
        }
    }
    *cost = std::sqrt(s2);
    return IWBC_OK;
}

int iwbc_qp_solve_batch(int device, int B, int n, int me, int mi, const double* H, const double* g, const double* CE, const double* ce0,
    const double* CI, const double* ci0, double* x, int* status, int* iters, unsigned* active)
{
    if (B <= 0 || n <= 0 || me < 0 || mi < 0 || !H || !g || !x || !status || (me && (!CE || !ce0)) || (mi && (!CI || !ci0))) {
        set_error("iwbc_qp_solve_batch: bad argument");
        return IWBC_ERR_ARG;
    }
    if (mi > 32 * IWBC_ACTIVE_WORDS || n > 128) {
        set_error("iwbc_qp_solve_batch: n <= 128 and mi <= 384 supported");
        return IWBC_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        set_error("no usable CUDA device: the batched QP has no CPU fallback");
        return IWBC_ERR_CUDA;
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    dev::Shape S;
    std::memset(&S, 0, sizeof(S));
    S.n = n, S.me = me, S.mi = mi, S.ldJ = n | 1, S.ldM = 1;
    dev::Sizes z;
    std::memset(&z, 0, sizeof(z));
    z.n = n, z.mi = mi, z.ldJ = S.ldJ, z.ldM = 1;
    dev::make_layout(z, S);
    const size_t Bz = B;
    double *dH, *dg, *dCE, *dce, *dCI, *dci, *dx;
    int *dst, *dit;
    unsigned* dact;
    CUDA_TRY(cudaMalloc(&dH, Bz * n * n * 8));
    CUDA_TRY(cudaMalloc(&dg, Bz * n * 8));
    CUDA_TRY(cudaMalloc(&dCE, std::max<size_t>(Bz * me * n, 1) * 8));
    CUDA_TRY(cudaMalloc(&dce, std::max<size_t>(Bz * me, 1) * 8));
    CUDA_TRY(cudaMalloc(&dCI, std::max<size_t>(Bz * mi * n, 1) * 8));
    CUDA_TRY(cudaMalloc(&dci, std::max<size_t>(Bz * mi, 1) * 8));
    CUDA_TRY(cudaMalloc(&dx, Bz * n * 8));
    CUDA_TRY(cudaMalloc(&dst, Bz * 4));
    CUDA_TRY(cudaMalloc(&dit, Bz * 4));
    CUDA_TRY(cudaMalloc(&dact, Bz * IWBC_ACTIVE_WORDS * 4));
    CUDA_TRY(cudaMemcpy(dH, H, Bz * n * n * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dg, g, Bz * n * 8, cudaMemcpyHostToDevice));
    if (me) {
        CUDA_TRY(cudaMemcpy(dCE, CE, Bz * me * n * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dce, ce0, Bz * me * 8, cudaMemcpyHostToDevice));
    }
    if (mi) {
        CUDA_TRY(cudaMemcpy(dCI, CI, Bz * mi * n * 8, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(dci, ci0, Bz * mi * 8, cudaMemcpyHostToDevice));
    }
    dev::QpIO io = {B, n, me, mi, dH, dg, dCE, dce, dCI, dci, dx, dst, dit, dact};
    int rc = IWBC_OK;
    if (n <= 32)
        rc = iwbc_launch_qp_dense_32(S, io, device, prop.multiProcessorCount);
    else if (n <= 64)
        rc = iwbc_launch_qp_dense_64(S, io, device, prop.multiProcessorCount);
    else if (n <= 96)
        rc = iwbc_launch_qp_dense_96(S, io, device, prop.multiProcessorCount);
    else
        rc = iwbc_launch_qp_dense_128(S, io, device, prop.multiProcessorCount);
    if (rc == IWBC_OK) {
        CUDA_TRY(cudaMemcpy(x, dx, Bz * n * 8, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(status, dst, Bz * 4, cudaMemcpyDeviceToHost));
        if (iters)
            CUDA_TRY(cudaMemcpy(iters, dit, Bz * 4, cudaMemcpyDeviceToHost));
        if (active)
            CUDA_TRY(cudaMemcpy(active, dact, Bz * IWBC_ACTIVE_WORDS * 4, cudaMemcpyDeviceToHost));
    }
    cudaFree(dH), cudaFree(dg), cudaFree(dCE), cudaFree(dce), cudaFree(dCI), cudaFree(dci), cudaFree(dx), cudaFree(dst), cudaFree(dit), cudaFree(dact);
    return rc;
}

} // extern "C"
