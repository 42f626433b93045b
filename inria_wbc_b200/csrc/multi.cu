// multi.cu -- one host call for a batch spread over several GPUs of one node (SURVEY.md 7 "shard/gather", 8(e)).
//
// Robot states are independent, so a batch of B states shards into contiguous ranges [g B / G, (g + 1) B / G), one per device;
// there is no collective on the data path (BASELINE.json north_star: "only host gather of torques and accelerations happens").
// iwbc_multi owns one iwbc_batch and one worker thread per device; iwbc_multi_update wakes the workers, every worker runs the
// host-buffer update of its own shard (its own upload / compute / download streams: iwbc_batch_update's pipeline) straight
// from / into the caller's arrays -- disjoint slices of the same buffers, which is the gather -- and the call returns when
// all of them are done.  With pinned host arrays (iwbc_host_alloc) the eight PCIe links run concurrently.
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "batch_state.hpp"

using namespace iwbc;

struct iwbc_multi {
    const iwbc_problem* prob = nullptr;
    std::vector<int> devices;
    std::vector<iwbc_batch*> batches;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    long generation = 0; // bumped per job
    int pending = 0;
    bool stop = false;
    // the current job
    int B = 0;
    const double *q = nullptr, *v = nullptr, *refs = nullptr;
    long refs_stride = 0;
    const unsigned* mask = nullptr;
    const iwbc_results_t* out = nullptr;
    std::vector<int> rc;
    std::vector<std::string> err;
};

extern "C" int iwbc_shard_range(int B, int nshards, int g, int* first, int* count)
{
    if (B < 0 || nshards <= 0 || g < 0 || g >= nshards)
        return IWBC_ERR_ARG;
    const long lo = (long)B * g / nshards, hi = (long)B * (g + 1) / nshards;
    if (first)
        *first = (int)lo;
    if (count)
        *count = (int)(hi - lo);
    return IWBC_OK;
}

namespace {
void worker_main(iwbc_multi* m, int g)
{
    cudaSetDevice(m->devices[g]);
    long seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv_job.wait(lk, [&] { return m->stop || m->generation != seen; });
            if (m->stop)
                return;
            seen = m->generation;
        }
        int first = 0, count = 0;
        iwbc_shard_range(m->B, (int)m->devices.size(), g, &first, &count);
        int rc = IWBC_OK;
        if (count > 0) {
            const iwbc_problem_t& P = m->prob->cs.prob;
            const size_t f = (size_t)first;
            const iwbc_results_t& o = *m->out;
            iwbc_results_t r;
            r.tau = o.tau ? o.tau + f * P.na : nullptr, r.dv = o.dv ? o.dv + f * P.nv : nullptr;
            r.f = o.f ? o.f + f * 12 * P.ncontacts : nullptr, r.q_next = o.q_next ? o.q_next + f * P.nq : nullptr;
            r.v_next = o.v_next ? o.v_next + f * P.nv : nullptr, r.q_cmd = o.q_cmd ? o.q_cmd + f * P.nv : nullptr;
            r.momentum = o.momentum ? o.momentum + f * 3 : nullptr, r.status = o.status ? o.status + f : nullptr;
            r.iters = o.iters ? o.iters + f : nullptr, r.active = o.active ? o.active + f * IWBC_ACTIVE_WORDS : nullptr;
            rc = iwbc_batch_update(m->batches[g], count, m->q + f * P.nq, m->v + f * P.nv, m->refs + (m->refs_stride ? f * (size_t)P.nref : 0), m->refs_stride,
                m->mask ? m->mask + f : nullptr, &r);
            if (rc)
                m->err[g] = iwbc_last_error(); // (thread-local on the worker: handed to the caller below)
        }
        std::lock_guard<std::mutex> lk(m->mu);
        m->rc[g] = rc;
        if (--m->pending == 0)
            m->cv_done.notify_all();
    }
}
} // namespace

extern "C" {

iwbc_multi* iwbc_multi_create(const iwbc_problem* p, int ndev, const int* devices, int max_batch_total)
{
    if (!p || ndev <= 0 || max_batch_total <= 0) {
        set_error("iwbc_multi_create: bad argument");
        return nullptr;
    }
    auto* m = new iwbc_multi;
    m->prob = p;
    for (int g = 0; g < ndev; ++g)
        m->devices.push_back(devices ? devices[g] : g);
    m->rc.assign(ndev, 0);
    m->err.assign(ndev, "");
    const int per = (max_batch_total + ndev - 1) / ndev;
    for (int g = 0; g < ndev; ++g) {
        iwbc_batch* b = iwbc_batch_create(p, m->devices[g], per);
        if (!b) {
            iwbc_multi_free(m);
            return nullptr;
        }
        m->batches.push_back(b);
    }
    for (int g = 0; g < ndev; ++g)
        m->workers.emplace_back(worker_main, m, g);
    return m;
}

void iwbc_multi_free(iwbc_multi* m)
{
    if (!m)
        return;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->stop = true;
    }
    m->cv_job.notify_all();
    for (auto& t : m->workers)
        t.join();
    for (iwbc_batch* b : m->batches)
        iwbc_batch_free(b);
    delete m;
}

int iwbc_multi_device_count(const iwbc_multi* m) { return m ? (int)m->devices.size() : 0; }

int iwbc_multi_update(iwbc_multi* m, int B, const double* q, const double* v, const double* refs, long refs_stride, const unsigned* contact_mask,
    const iwbc_results_t* out)
{
    if (!m || !q || !v || !refs || !out || B < 0) {
        set_error("iwbc_multi_update: bad argument");
        return IWBC_ERR_ARG;
    }
    const int G = (int)m->devices.size();
    for (int g = 0; g < G; ++g) {
        int first, count;
        iwbc_shard_range(B, G, g, &first, &count);
        if (count > m->batches[g]->max_batch) {
            set_error("iwbc_multi_update: B = " + std::to_string(B) + " exceeds the capacity given at creation");
            return IWBC_ERR_ARG;
        }
    }
    if (B == 0)
        return IWBC_OK;
    {
        std::unique_lock<std::mutex> lk(m->mu);
        m->B = B, m->q = q, m->v = v, m->refs = refs, m->refs_stride = refs_stride, m->mask = contact_mask, m->out = out;
        m->pending = G;
        m->generation++;
        m->cv_job.notify_all();
        m->cv_done.wait(lk, [&] { return m->pending == 0; });
    }
    for (int g = 0; g < G; ++g)
        if (m->rc[g]) {
            set_error("device " + std::to_string(m->devices[g]) + ": " + m->err[g]);
            return m->rc[g];
        }
    return IWBC_OK;
}

long iwbc_multi_launch_count(const iwbc_multi* m)
{
    long n = 0;
    if (m)
        for (iwbc_batch* b : m->batches)
            n += b->launches;
    return n;
}

void* iwbc_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        set_error("iwbc_host_alloc: cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}

void iwbc_host_free(void* p)
{
    if (p)
        cudaFreeHost(p);
}

} // extern "C"
