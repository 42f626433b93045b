// batch_state.hpp -- the opaque handle behind `iwbc_batch` and the launch entry points of the per-shape
// translation units (fast_tu.cuh); shared by batch.cu and fast_*.cu.
#pragma once
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/iwbc_batch.h"
#include "cycle_kernel.cuh"
#include "fast_kernel.cuh"
#include "host_state.hpp"

#define CUDA_TRY(expr)                                                                                        \
    do {                                                                                                      \
        cudaError_t e_ = (expr);                                                                              \
        if (e_ != cudaSuccess) {                                                                              \
            iwbc::set_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
            return IWBC_ERR_CUDA;                                                                             \
        }                                                                                                     \
    } while (0)

// Everything that depends on (problem revision, contact mask): the QP shape and shared-memory layout (passed to the kernels as
// __grid_constant__ parameters), the fast path's plan (its own device buffer) and the launch shapes.  Cached per mask in the
// batch, so a walking tick that alternates between three masks neither re-derives nor re-uploads anything.
struct iwbc_mask_state {
    unsigned mask = ~0u;
    unsigned long revision = ~0ul;
    iwbc::dev::Shape shape;
    iwbc::dev::FShape fshape;
    iwbc::dev::FPlan* d_fplan = nullptr;
    int G = 0, IPC = 0, grid = 0, fast_grid = 0, slow_grid = 0;
    size_t smem_bytes = 0, fast_smem = 0, slow_smem = 0;
    bool fast_ok = false;
};

struct iwbc_batch {
    const iwbc_problem* prob = nullptr;
    unsigned long revision = ~0ul;
    std::vector<iwbc_mask_state> mask_cache; // the inactive masks' states (the active one lives in the fields below)
    unsigned long dprob_revision = ~0ul; // revision of the problem held by d_prob
    int device = 0;
    int max_batch = 0;
    int sm_count = 0;
    iwbc_problem_t* d_prob = nullptr;
    cudaStream_t stream = nullptr;
    // staging buffers (device)
    double *d_q = nullptr, *d_v = nullptr, *d_refs = nullptr;
    double *d_tau = nullptr, *d_dv = nullptr, *d_f = nullptr, *d_qn = nullptr, *d_vn = nullptr, *d_qc = nullptr, *d_mom = nullptr;
    int *d_status = nullptr, *d_iters = nullptr;
    unsigned* d_active = nullptr;
    double* d_sensors = nullptr;  // iwbc_batch_update_sensors: [max_batch][15] staging (allocated on first use)
    bool sensor_call = false;     // set while iwbc_batch_update_sensors_device drives the cycle (the plain entry points refuse an activated stabilizer)
    int* d_phase = nullptr;       // iwbc_batch_behavior_run: per-instance phase offsets, contact masks (allocated on first use)
    unsigned* d_cmask = nullptr;
    // bucketing scratch (host)
    std::vector<int> perm;
    std::vector<double> hq, hv, hr;
    std::vector<char> stage;
    // launch bookkeeping
    long launches = 0;
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double ms_sum = 0;
    long ms_n = 0;
    unsigned shape_mask = ~0u;
    unsigned long shape_revision = ~0ul; // problem revision the current shape / plan were derived from
    iwbc::dev::Shape shape;
    int G = 0, IPC = 0;
    long long last_cycles[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long* d_prof = nullptr; // steady-state phase profile accumulators (enabled on demand): [0..6] active-set kernel, [8..14] fast kernel
    bool profiling = false;
    unsigned long long loop_prof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; // active-set loop phases read by the last iwbc_batch_phase_profile
    // fast path (fast_kernel.cuh): equality-constrained first pass, 3 CTAs / SM; the active-set kernel continues the rest
    bool fast_ok = false, fast_enabled = true;
    iwbc::dev::FShape fshape;
    iwbc::dev::FPlan fplan;
    iwbc::dev::FPlan* d_fplan = nullptr;
    int *d_slow_list = nullptr, *d_slow_count = nullptr; // per launch slot: list region [list_off, ...), counter d_slow_count[slot]
    int slot = 0, list_off = 0;
    double* d_scratch = nullptr; // front-end images handed from k_fast to k_slow
    int* d_slow_code = nullptr;
    int scratch_cap = 0, scratch_off = 0, scratch_len = 0, scratch_img = 0; // capacity (instances); sub-range used by the current launch
    cudaStream_t aux[2] = {nullptr, nullptr}; // copy / compute pipelining of iwbc_batch_update: two compute streams (kernel tails of successive chunks overlap),
    cudaStream_t s_in = nullptr, s_out = nullptr; // ... one upload stream, one download stream,
    cudaEvent_t ev_in[16] = {}, ev_done[16] = {}; // ... chained per chunk by events
    cudaEvent_t ev_const = nullptr;
    int fast_grid = 0, slow_grid = 0;
    size_t fast_smem = 0, slow_smem = 0;
    bool plan_dirty = true;
    size_t smem_bytes = 0;
    int grid = 0;
};


constexpr int IWBC_FAST_THREADS = 128; // threads per CTA of k_fast / k_slow

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and per device, not per batch: every launch site keeps the largest
// value it has set (per device) and only ever raises it, under a lock (two host threads may drive two handles on one device)
int iwbc_raise_smem_limit(const void* func, int device, size_t bytes, size_t* max_set /* [16] per launch site */);

// k_fast / k_slow launch pair of one compile-time (nv, me) shape
int iwbc_launch_fast_50_18(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);
int iwbc_launch_fast_50_12(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);
int iwbc_launch_fast_38_18(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);
int iwbc_launch_fast_38_12(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);
int iwbc_launch_fast_0_0(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);

// generic cycle kernel per group width, its thread-per-instance form per working-set size, the dense-QP kernel (cycle_tu.cuh)
namespace iwbc {
namespace dev {
    struct QpIO { // H[B][n][n] etc. in global memory: the SolverHQPBase-shaped entry point
        int B, n, me, mi;
        const double *H, *g, *CE, *ce0, *CI, *ci0;
        double* x;
        int* status;
        int* iters;
        unsigned* active;
    };
}
}
int iwbc_launch_cycle_32_4_3(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);
int iwbc_launch_cycle_32_4_4(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);
int iwbc_launch_cycle_64(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);
int iwbc_launch_cycle_96(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);
int iwbc_launch_cycle_128(iwbc_batch* b, const iwbc::dev::IO& io, cudaStream_t st);
int iwbc_launch_cycle_tpi_1024(iwbc_batch* b, const iwbc::dev::IO& io, int grid, cudaStream_t st);
int iwbc_launch_cycle_tpi_1792(iwbc_batch* b, const iwbc::dev::IO& io, int grid, cudaStream_t st);
int iwbc_launch_qp_dense_32(const iwbc::dev::Shape& S, const iwbc::dev::QpIO& io, int device, int sm_count);
int iwbc_launch_qp_dense_64(const iwbc::dev::Shape& S, const iwbc::dev::QpIO& io, int device, int sm_count);
int iwbc_launch_qp_dense_96(const iwbc::dev::Shape& S, const iwbc::dev::QpIO& io, int device, int sm_count);
int iwbc_launch_qp_dense_128(const iwbc::dev::Shape& S, const iwbc::dev::QpIO& io, int device, int sm_count);

// one pass of the control cycle over io's instances (io.list / io.list_count: optional device-side instance list) with one
// contact mask: prepares the shape, refreshes constants, launches (batch.cu)
int iwbc_run_cycle(iwbc_batch* b, const iwbc::dev::IO& io, unsigned mask, cudaStream_t st);

struct iwbc_collision;
