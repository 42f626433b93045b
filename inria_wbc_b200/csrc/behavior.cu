// behavior.cu -- device-side behaviours: the reference generators that call the hot path, evaluated on the
// GPU for every (instance, tick) from a compiled segment table (behavior.hpp), and the closed device-side roll-out
//     for tick: refs, contact masks  <- k_behavior_refs          (replaces Behavior::update's setters, e.g. walk.cpp:198-229)
//               per-mask instance lists (device-side bucketing, no host round trip)
//               one shape-uniform launch set of the control cycle per contact mask, state integrated in place
#include <algorithm>
#include <cstring>
#include <string>

#include "batch_state.hpp"
#include "behavior.hpp"

using namespace iwbc;

namespace {

struct BehDev {
    int nseg, nch, total_ticks, loop, loop_start, nref, nmasks;
    double dt;
    const iwbc_beh_channel_t* channels;
    const iwbc_beh_seg_t* segs;
    const iwbc_beh_chanseg_t* cs;
    const double* ref0;
    const double* table;
    unsigned masks[8];
};

// minimum_jerk_polynom<ORDER> (include/inria_wbc/trajs/trajectory_generator.hpp:31-60), scalar factor of (xf - x0)
__device__ __forceinline__ double min_jerk_factor(double t, double duration, int order)
{
    const double td = t / duration;
    if (order == 0) {
        const double td3 = td * td * td;
        return 6 * td * td * td3 - 15 * td * td3 + 10 * td3;
    }
    if (order == 1) {
        const double td2 = td * td;
        return (30 * td * td * td2 - 60 * td * td2 + 30 * td2) / duration;
    }
    return (120 * td * td * td - 180 * td * td + 60 * td) / (duration * duration);
}

// refs[i][:] <- construction-time references (once per roll-out)
__global__ void k_behavior_init(BehDev d, int B, double* refs)
{
    const long total = (long)B * d.nref;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
        refs[e] = d.ref0[e % d.nref];
}

// One thread per (instance, channel): the sample the behaviour writes before the controller update of tick `tick`
// (+ phase[i]); channel 0's thread also publishes the instance's contact mask and appends it to its mask's bucket.
__global__ void k_behavior_refs(BehDev d, int B, const int* __restrict__ phase, int tick, double* __restrict__ refs, unsigned* __restrict__ mask_out,
    int* __restrict__ bcount, int* __restrict__ blist)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = gid / d.nch, c = gid - i * d.nch;
    if (i >= B)
        return;
    long g = (long)tick + (phase ? phase[i] : 0);
    if (d.loop) {
        if (g >= d.total_ticks)
            g = d.loop_start + (g - d.loop_start) % (d.total_ticks - d.loop_start);
    }
    else if (g >= d.total_ticks)
        g = d.total_ticks - 1; // run_ = false: the last written references stay (walk.cpp:236-243)
    int s = 0;
    while (s + 1 < d.nseg && g >= d.segs[s + 1].start)
        ++s;
    const iwbc_beh_seg_t sg = d.segs[s];
    const int local = (int)g - sg.start;
    const iwbc_beh_channel_t ch = d.channels[c];
    const iwbc_beh_chanseg_t& cs = d.cs[s * d.nch + c];
    if (ch.table_rows > 0) {
        // recorded trajectory, forwards then backwards (cartesian_traj.cpp:38-44); set_se3_ref(SE3) / set_com_ref(Vector3): zero derivatives
        const int t = local < ch.table_rows ? local : 2 * ch.table_rows - 2 - local;
        const int width = ch.kind == 0 ? 3 : 12;
        const double* row = d.table + ch.table_off + (size_t)t * width;
        double* o = refs + (size_t)i * d.nref + ch.slot[0];
        for (int k = 0; k < 3; ++k)
            o[k] = row[k] * ch.scale;
        if (ch.kind == 0) {
            for (int k = 3; k < 9; ++k)
                o[k] = 0.0;
        }
        else {
            for (int k = 3; k < 12; ++k)
                o[k] = row[k];
            for (int k = 12; k < 24; ++k)
                o[k] = 0.0;
        }
    }
    else {
    double p[3], vel[3] = {0, 0, 0}, acc[3] = {0, 0, 0};
    double f0 = 0, f1 = 0, f2 = 0;
    if (cs.minjerk) {
        const double t = d.dt * (double)local;
        f0 = min_jerk_factor(t, cs.duration, 0);
        if (ch.ff) {
            f1 = min_jerk_factor(t, cs.duration, 1);
            f2 = min_jerk_factor(t, cs.duration, 2);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double dx = cs.p1[k] - cs.p0[k];
        p[k] = cs.minjerk ? cs.p0[k] + dx * f0 : cs.p0[k];
        vel[k] = dx * f1;
        acc[k] = dx * f2;
    }
    for (int w = 0; w < 2; ++w) {
        if (ch.slot[w] < 0)
            continue;
        double* o = refs + (size_t)i * d.nref + ch.slot[w];
        if (ch.kind == 0) { // TrajectorySample(3,3): pos vel acc
#pragma unroll
            for (int k = 0; k < 3; ++k)
                o[k] = p[k], o[3 + k] = vel[k], o[6 + k] = acc[k];
        }
        else { // TrajectorySample(12,6): p, R (column-major), vel(6), acc(6)
#pragma unroll
            for (int k = 0; k < 3; ++k)
                o[k] = p[k];
            double wv[3] = {0, 0, 0}, wa[3] = {0, 0, 0};
            if (cs.rot) {
                // R = R0 * AngleAxis(angle * f0, axis); angular parts of vel / acc = R0 * (d angle * axis)  (trajectory_generator.hpp:96-160)
                const double a = cs.angle * f0, sa = sin(a), ca = cos(a), ux = cs.axis[0], uy = cs.axis[1], uz = cs.axis[2], oc = 1.0 - ca;
                const double E[9] = {ca + ux * ux * oc, ux * uy * oc - uz * sa, ux * uz * oc + uy * sa, uy * ux * oc + uz * sa, ca + uy * uy * oc,
                    uy * uz * oc - ux * sa, uz * ux * oc - uy * sa, uz * uy * oc + ux * sa, ca + uz * uz * oc}; // row-major
                // R0 is held column-major: R0(r, k) = cs.R0[3 k + r]
#pragma unroll
                for (int col = 0; col < 3; ++col)
#pragma unroll
                    for (int r = 0; r < 3; ++r)
                        o[3 + 3 * col + r] = cs.R0[r] * E[col] + cs.R0[3 + r] * E[3 + col] + cs.R0[6 + r] * E[6 + col];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double ra = cs.R0[r] * ux + cs.R0[3 + r] * uy + cs.R0[6 + r] * uz;
                    wv[r] = ra * (cs.angle * f1);
                    wa[r] = ra * (cs.angle * f2);
                }
            }
            else {
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    o[3 + k] = cs.R0[k];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                o[12 + k] = vel[k], o[15 + k] = wv[k];
                o[18 + k] = acc[k], o[21 + k] = wa[k];
            }
        }
    }
    }
    if (c == 0) {
        const unsigned m = (local == sg.ticks - 1) ? sg.mask_last : sg.mask;
        if (mask_out)
            mask_out[i] = m;
        if (bcount) {
            int k = 0;
            while (k + 1 < d.nmasks && d.masks[k] != m)
                ++k;
            const int pos = atomicAdd(bcount + k, 1); // order inside a bucket is arbitrary; no result depends on it
            blist[(size_t)k * B + pos] = i;
        }
    }
}

int ensure_device(iwbc_behavior* h, int device, int B)
{
    CUDA_TRY(cudaSetDevice(device));
    if (h->device != device) {
        if (h->d_blob) {
            set_error("iwbc_behavior: one behaviour handle serves one device");
            return IWBC_ERR_ARG;
        }
        const size_t s_ch = h->channels.size() * sizeof(iwbc_beh_channel_t), s_sg = h->segs.size() * sizeof(iwbc_beh_seg_t),
                     s_cs = h->cs.size() * sizeof(iwbc_beh_chanseg_t), s_r0 = h->prob->cs.ref0.size() * sizeof(double),
                     s_tb = h->table.size() * sizeof(double);
        auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
        char* blob = nullptr;
        CUDA_TRY(cudaMalloc((void**)&blob, up(s_ch) + up(s_sg) + up(s_cs) + up(s_r0) + up(s_tb) + 256));
        h->d_blob = blob;
        h->d_channels = (iwbc_beh_channel_t*)blob, blob += up(s_ch);
        h->d_segs = (iwbc_beh_seg_t*)blob, blob += up(s_sg);
        h->d_cs = (iwbc_beh_chanseg_t*)blob, blob += up(s_cs);
        h->d_ref0 = (double*)blob, blob += up(s_r0);
        h->d_table = (double*)blob, blob += up(s_tb);
        h->d_bcount = (int*)blob;
        CUDA_TRY(cudaMemcpy(h->d_channels, h->channels.data(), s_ch, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->d_segs, h->segs.data(), s_sg, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->d_cs, h->cs.data(), s_cs, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(h->d_ref0, h->prob->cs.ref0.data(), s_r0, cudaMemcpyHostToDevice));
        if (s_tb)
            CUDA_TRY(cudaMemcpy(h->d_table, h->table.data(), s_tb, cudaMemcpyHostToDevice));
        h->device = device;
    }
    if (B > h->blist_cap) {
        if (h->d_blist) {
            CUDA_TRY(cudaDeviceSynchronize());
            cudaFree(h->d_blist);
        }
        CUDA_TRY(cudaMalloc((void**)&h->d_blist, (size_t)B * h->masks.size() * sizeof(int)));
        h->blist_cap = B;
    }
    return IWBC_OK;
}

BehDev dev_view(const iwbc_behavior* h)
{
    BehDev d;
    std::memset(&d, 0, sizeof(d));
    d.nseg = (int)h->segs.size(), d.nch = (int)h->channels.size(), d.total_ticks = h->total_ticks, d.loop = h->loop ? 1 : 0, d.loop_start = h->loop_start;
    d.nref = (int)h->prob->cs.ref0.size(), d.nmasks = (int)h->masks.size(), d.dt = h->dt;
    d.channels = h->d_channels, d.segs = h->d_segs, d.cs = h->d_cs, d.ref0 = h->d_ref0, d.table = h->d_table;
    for (size_t k = 0; k < h->masks.size(); ++k)
        d.masks[k] = h->masks[k];
    return d;
}

int eval_tick(iwbc_behavior* h, int B, const int* phase, int tick, double* refs, unsigned* mask_out, bool buckets, cudaStream_t st)
{
    const BehDev d = dev_view(h);
    if (buckets)
        CUDA_TRY(cudaMemsetAsync(h->d_bcount, 0, 8 * sizeof(int), st));
    const long threads = (long)B * d.nch;
    k_behavior_refs<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(d, B, phase, tick, refs, mask_out, buckets ? h->d_bcount : nullptr,
        buckets ? h->d_blist : nullptr);
    CUDA_TRY(cudaGetLastError());
    return IWBC_OK;
}

} // namespace

extern "C" {

iwbc_behavior* iwbc_behavior_load(iwbc_problem* p, const char* behavior_yaml, const char* overrides_yaml)
{
    if (!p || !behavior_yaml) {
        set_error("iwbc_behavior_load: NULL argument");
        return nullptr;
    }
    iwbc_behavior* h = new iwbc_behavior();
    try {
        host::compile_behavior(*h, *p, behavior_yaml, overrides_yaml ? overrides_yaml : "");
    }
    catch (std::exception& e) {
        set_error(e.what());
        delete h;
        return nullptr;
    }
    return h;
}

void iwbc_behavior_free(iwbc_behavior* h)
{
    if (!h)
        return;
    if (h->d_blob || h->d_blist) {
        cudaSetDevice(h->device);
        cudaFree(h->d_blob);
        cudaFree(h->d_blist);
    }
    delete h;
}

int iwbc_behavior_info(const iwbc_behavior* h, int* total_ticks, int* nsegments, int* nchannels, int* nmasks, unsigned* masks8)
{
    if (!h)
        return IWBC_ERR_ARG;
    if (total_ticks)
        *total_ticks = h->total_ticks;
    if (nsegments)
        *nsegments = (int)h->segs.size();
    if (nchannels)
        *nchannels = (int)h->channels.size();
    if (nmasks)
        *nmasks = (int)h->masks.size();
    if (masks8)
        for (size_t k = 0; k < h->masks.size(); ++k)
            masks8[k] = h->masks[k];
    return IWBC_OK;
}

int iwbc_behavior_segment(const iwbc_behavior* h, int seg, int* start, int* ticks, unsigned* mask, unsigned* mask_last, double* waypoints)
{
    if (!h || seg < 0 || seg >= (int)h->segs.size()) {
        set_error("iwbc_behavior_segment: bad segment index");
        return IWBC_ERR_ARG;
    }
    const iwbc_beh_seg_t& s = h->segs[seg];
    if (start)
        *start = s.start;
    if (ticks)
        *ticks = s.ticks;
    if (mask)
        *mask = s.mask;
    if (mask_last)
        *mask_last = s.mask_last;
    if (waypoints) {
        const int nch = (int)h->channels.size();
        for (int c = 0; c < nch; ++c) {
            const iwbc_beh_chanseg_t& cs = h->cs[seg * nch + c];
            double* o = waypoints + 8 * c;
            for (int k = 0; k < 3; ++k)
                o[k] = cs.p0[k], o[3 + k] = cs.p1[k];
            o[6] = cs.duration;
            o[7] = cs.minjerk;
        }
    }
    return IWBC_OK;
}

int iwbc_behavior_eval_device(iwbc_behavior* h, int device, int B, const int* phase, int tick, double* refs, unsigned* contact_mask, int init_refs,
    void* stream)
{
    if (!h || !refs || B < 0) {
        set_error("iwbc_behavior_eval_device: bad argument");
        return IWBC_ERR_ARG;
    }
    if (B == 0)
        return IWBC_OK;
    int rc = ensure_device(h, device, B);
    if (rc)
        return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (init_refs) {
        k_behavior_init<<<1184, 256, 0, st>>>(dev_view(h), B, refs);
        CUDA_TRY(cudaGetLastError());
    }
    return eval_tick(h, B, phase, tick, refs, contact_mask, false, st);
}

static int behavior_rollout(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, double* refs, const int* phase, int tick0, int ticks,
    int init_refs, const iwbc_results_t* out_last, unsigned* contact_mask, double* q_cmd_trace, void* stream)
{
    if (!b || !h || !q || !v || !refs || B < 0 || ticks < 0) {
        set_error("iwbc_batch_behavior_rollout_device: bad argument");
        return IWBC_ERR_ARG;
    }
    if (h->prob != b->prob) {
        set_error("iwbc_batch_behavior_rollout_device: the behaviour was compiled for another controller");
        return IWBC_ERR_ARG;
    }
    if (B > b->max_batch) {
        set_error("iwbc_batch_behavior_rollout_device: B exceeds max_batch");
        return IWBC_ERR_ARG;
    }
    if (B == 0 || ticks == 0)
        return IWBC_OK;
    int rc = ensure_device(h, b->device, B);
    if (rc)
        return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
    const int nref = (int)h->prob->cs.ref0.size();
    if (init_refs) {
        k_behavior_init<<<1184, 256, 0, st>>>(dev_view(h), B, refs);
        CUDA_TRY(cudaGetLastError());
    }
    for (int t = 0; t < ticks; ++t) {
        rc = eval_tick(h, B, phase, tick0 + t, refs, contact_mask, true, st);
        if (rc)
            return rc;
        b->launches++;
        for (size_t k = 0; k < h->masks.size(); ++k) {
            dev::IO io;
            std::memset(&io, 0, sizeof(io));
            io.B = B;
            io.q = q, io.v = v, io.refs = refs, io.refs_stride = nref;
            io.q_rw = q, io.v_rw = v;
            io.sticky_status = t > 0; // a failed tick freezes the instance and its status survives the call (batch.cu, roll-out)
            io.in_list = h->d_blist + (size_t)k * B;
            io.in_count = h->d_bcount + k;
            if (out_last) {
                io.status = out_last->status, io.iters = out_last->iters;
                if (t == ticks - 1) {
                    io.tau = out_last->tau, io.dv = out_last->dv, io.f = out_last->f, io.q_cmd = out_last->q_cmd;
                    io.momentum = out_last->momentum, io.active = out_last->active, io.q_next = out_last->q_next, io.v_next = out_last->v_next;
                }
            }
            if (q_cmd_trace) // every tick's command vector (what the reference's tests record with controller->q(false) after each update)
                io.q_cmd = q_cmd_trace + (size_t)t * B * h->prob->cs.prob.nv;
            rc = iwbc_run_cycle(b, io, h->masks[k], st);
            if (rc)
                return rc;
        }
    }
    if (q_cmd_trace && out_last && out_last->q_cmd)
        CUDA_TRY(cudaMemcpyAsync(out_last->q_cmd, q_cmd_trace + (size_t)(ticks - 1) * B * h->prob->cs.prob.nv, (size_t)B * h->prob->cs.prob.nv * 8,
            cudaMemcpyDeviceToDevice, st));
    return IWBC_OK;
}

int iwbc_batch_behavior_rollout_device(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, double* refs, const int* phase, int tick0, int ticks,
    int init_refs, const iwbc_results_t* out_last, unsigned* contact_mask, void* stream)
{
    return behavior_rollout(b, h, B, q, v, refs, phase, tick0, ticks, init_refs, out_last, contact_mask, nullptr, stream);
}

int iwbc_batch_behavior_rollout_trace_device(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, double* refs, const int* phase, int tick0,
    int ticks, int init_refs, const iwbc_results_t* out_last, unsigned* contact_mask, double* q_cmd_trace, void* stream)
{
    if (!q_cmd_trace) {
        set_error("iwbc_batch_behavior_rollout_trace_device: NULL trace");
        return IWBC_ERR_ARG;
    }
    return behavior_rollout(b, h, B, q, v, refs, phase, tick0, ticks, init_refs, out_last, contact_mask, q_cmd_trace, stream);
}

// Host-buffer form of the roll-out (what a `for (tick) behavior->update();` loop of the reference does, src/robot_dart/talos.cpp:380-390):
// uploads the state once, runs `ticks` closed device-side ticks, downloads the advanced state and the last tick's results.
static int behavior_run(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, const int* phase, int tick0, int ticks,
    const iwbc_results_t* out_last, unsigned* contact_mask, double* q_cmd_trace)
{
    if (!b || !h || !q || !v || B < 0 || ticks < 0) {
        set_error("iwbc_batch_behavior_run: bad argument");
        return IWBC_ERR_ARG;
    }
    if (B > b->max_batch) {
        set_error("iwbc_batch_behavior_run: B exceeds max_batch");
        return IWBC_ERR_ARG;
    }
    if (B == 0 || ticks == 0)
        return IWBC_OK;
    CUDA_TRY(cudaSetDevice(b->device));
    const iwbc_problem_t& P = b->prob->cs.prob;
    cudaStream_t st = b->stream;
    const size_t nq = P.nq, nv = P.nv, na = P.na, nf = 12 * P.ncontacts;
    if (!b->d_phase) {
        CUDA_TRY(cudaMalloc((void**)&b->d_phase, (size_t)b->max_batch * sizeof(int)));
        CUDA_TRY(cudaMalloc((void**)&b->d_cmask, (size_t)b->max_batch * sizeof(unsigned)));
    }
    CUDA_TRY(cudaMemcpyAsync(b->d_q, q, B * nq * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(b->d_v, v, B * nv * 8, cudaMemcpyHostToDevice, st));
    if (phase)
        CUDA_TRY(cudaMemcpyAsync(b->d_phase, phase, B * sizeof(int), cudaMemcpyHostToDevice, st));
    iwbc_results_t r;
    std::memset(&r, 0, sizeof(r));
    if (out_last) {
        r.tau = out_last->tau ? b->d_tau : nullptr, r.dv = out_last->dv ? b->d_dv : nullptr, r.f = out_last->f ? b->d_f : nullptr;
        r.q_cmd = out_last->q_cmd ? b->d_qc : nullptr, r.momentum = out_last->momentum ? b->d_mom : nullptr;
        r.status = out_last->status ? b->d_status : nullptr, r.iters = out_last->iters ? b->d_iters : nullptr;
        r.active = out_last->active ? b->d_active : nullptr;
    }
    // a fresh behaviour (tick0 == 0) starts from the construction-time references; later calls continue from the rows the previous call left
    int rc = IWBC_OK;
    if (!q_cmd_trace)
        rc = behavior_rollout(b, h, B, b->d_q, b->d_v, b->d_refs, phase ? b->d_phase : nullptr, tick0, ticks, tick0 == 0 ? 1 : 0, &r,
            contact_mask ? b->d_cmask : nullptr, nullptr, st);
    else {
        // per-tick command trace: device ring of <= 16 MB, drained to the host chunk by chunk
        const size_t row = (size_t)B * nv;
        const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)ticks, (16u << 20) / (row * 8)));
        double* d_trace = nullptr;
        CUDA_TRY(cudaMalloc((void**)&d_trace, (size_t)chunk * row * 8));
        for (int t0 = 0; t0 < ticks && rc == IWBC_OK; t0 += chunk) {
            const int n = std::min(chunk, ticks - t0);
            rc = behavior_rollout(b, h, B, b->d_q, b->d_v, b->d_refs, phase ? b->d_phase : nullptr, tick0 + t0, n, (tick0 == 0 && t0 == 0) ? 1 : 0, &r,
                contact_mask ? b->d_cmask : nullptr, d_trace, st);
            if (rc == IWBC_OK && cudaMemcpyAsync(q_cmd_trace + (size_t)t0 * row, d_trace, (size_t)n * row * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess)
                rc = IWBC_ERR_CUDA;
            if (rc == IWBC_OK && cudaStreamSynchronize(st) != cudaSuccess) // the ring is reused by the next chunk
                rc = IWBC_ERR_CUDA;
        }
        cudaFree(d_trace);
        if (rc == IWBC_ERR_CUDA && std::string(iwbc_last_error()).empty())
            set_error("iwbc_batch_behavior_run_trace: CUDA error");
    }
    if (rc)
        return rc;
    CUDA_TRY(cudaMemcpyAsync(q, b->d_q, B * nq * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(v, b->d_v, B * nv * 8, cudaMemcpyDeviceToHost, st));
    if (out_last) {
        auto back = [&](void* host, const void* devp, size_t bytes) -> int {
            if (host)
                CUDA_TRY(cudaMemcpyAsync(host, devp, bytes, cudaMemcpyDeviceToHost, st));
            return IWBC_OK;
        };
        rc = back(out_last->tau, b->d_tau, B * na * 8) | back(out_last->dv, b->d_dv, B * nv * 8) | back(out_last->f, b->d_f, B * nf * 8)
            | back(out_last->q_cmd, b->d_qc, B * nv * 8) | back(out_last->momentum, b->d_mom, B * 3 * 8) | back(out_last->status, b->d_status, B * 4)
            | back(out_last->iters, b->d_iters, B * 4) | back(out_last->active, b->d_active, (size_t)B * IWBC_ACTIVE_WORDS * 4);
        if (rc)
            return IWBC_ERR_CUDA;
        // q_next / v_next of the last tick ARE the advanced state
        if (out_last->q_next)
            CUDA_TRY(cudaMemcpyAsync(out_last->q_next, b->d_q, B * nq * 8, cudaMemcpyDeviceToHost, st));
        if (out_last->v_next)
            CUDA_TRY(cudaMemcpyAsync(out_last->v_next, b->d_v, B * nv * 8, cudaMemcpyDeviceToHost, st));
    }
    if (contact_mask)
        CUDA_TRY(cudaMemcpyAsync(contact_mask, b->d_cmask, B * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return IWBC_OK;
}

int iwbc_batch_behavior_run(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, const int* phase, int tick0, int ticks,
    const iwbc_results_t* out_last, unsigned* contact_mask)
{
    return behavior_run(b, h, B, q, v, phase, tick0, ticks, out_last, contact_mask, nullptr);
}

int iwbc_batch_behavior_run_trace(iwbc_batch* b, iwbc_behavior* h, int B, double* q, double* v, const int* phase, int tick0, int ticks,
    const iwbc_results_t* out_last, unsigned* contact_mask, double* q_cmd_trace)
{
    if (!q_cmd_trace) {
        set_error("iwbc_batch_behavior_run_trace: NULL trace");
        return IWBC_ERR_ARG;
    }
    return behavior_run(b, h, B, q, v, phase, tick0, ticks, out_last, contact_mask, q_cmd_trace);
}

} // extern "C"
