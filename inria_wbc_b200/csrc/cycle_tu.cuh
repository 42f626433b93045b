// cycle_tu.cuh -- the generic cycle kernel (one group of G threads per instance), its thread-per-instance form for large arm
// batches and the dense-QP kernel, one translation unit per group width so that the widths compile in parallel.
// Usage: #define IWBC_CYC_G 32|64|96|128 (group kernels + dense QP) or IWBC_CYC_TPI 1024|1792 (thread-per-instance working-set
// doubles), then #include "cycle_tu.cuh".
#include <algorithm>

#include "batch_state.hpp"

namespace iwbc {
namespace dev {

    // ------------------------------------------------------------------------------ the kernel
    // One group of G threads per instance, IPC instances per CTA, persistent over the batch.
    // (the warp-per-instance form is occupancy-bound by registers: three CTAs per SM need <= 168; without the bound ptxas drifts to
    // 216 and Franka drops from 38 M to 27 M updates/s)
    template <int G, int IPC, int MINB = (G <= 32 && G * IPC <= 128 ? 3 : 1)>
    __global__ void __launch_bounds__(G* IPC, MINB) k_cycle(const iwbc_problem_t* __restrict__ gP, const __grid_constant__ Shape S, IO io)
    {
        extern __shared__ double smem[];
        const int grp = threadIdx.x / G;
        const int tid = threadIdx.x - grp * G;
        Sizes z;
        z.nv = gP->nv;
        z.nq = gP->nq;
        z.nj = gP->nj;
        z.nf = gP->nframes;
        z.nc = gP->ncontacts;
        z.n = S.n;
        z.mi = S.mi;
        z.rows_dense = S.rows_dense;
        z.nref = gP->nref;
        z.ldJ = S.ldJ;
        z.ldM = S.ldM;
        z.S = &S;
        z.P = gP;
        const Sm sm = make_sm(smem + (size_t)grp * S.total, z);
        int flip = 0;

        // dof -> joint table
        for (int j = tid; j < z.nj; j += G) {
            const int c0 = gP->idx_v[j], cn = gP->jtype[j] == IWBC_J_FREEFLYER ? 6 : 1;
            for (int k = 0; k < cn; ++k)
                sm.dofj[c0 + k] = j;
        }
        const int ngroups = gridDim.x * IPC;
        const int count = io.list ? *io.list_count : io.B;
        if (io.work) {
            // dynamic hand-out: instances differ in active-set iterations (Tiago: 1 .. 30), a static stride leaves groups idle at the end
            __shared__ int s_next[IPC];
            for (;;) {
                int idx;
                if (G <= 32) {
                    idx = 0;
                    if (tid == 0)
                        idx = atomicAdd(io.work, 1);
                    idx = __shfl_sync(Grp<G>::mask(), idx, (threadIdx.x & 31) / G * G);
                }
                else {
                    if (tid == 0)
                        s_next[grp] = atomicAdd(io.work, 1);
                    __syncthreads();
                    idx = s_next[grp];
                }
                if (idx >= count)
                    break;
                const int inst = io.list ? io.list[idx] : idx;
                cycle_instance<G>(gP, sm, z, io, inst, tid, flip);
                Grp<G>::sync();
            }
            return;
        }
        for (int idx = blockIdx.x * IPC + grp; idx < count; idx += ngroups) {
            const int inst = io.list ? io.list[idx] : idx;
            cycle_instance<G>(gP, sm, z, io, inst, tid, flip);
            Grp<G>::sync();
        }
    }

    // ---- thread-per-instance form of the same cycle for large batches of the small arms: the generic device code
    // with a "group" of ONE thread (what tests/cpu_sim.cpp runs on the host), its working set in the thread's local memory
    // (interleaved across the warp by the hardware: lane-adjacent instances touch adjacent words).  A warp instruction then
    // advances 32 instances instead of one.
    template <int MAXD, int MINB>
    __global__ void __launch_bounds__(128, MINB) k_cycle_tpi(const iwbc_problem_t* __restrict__ gP, const __grid_constant__ Shape S, IO io, int first, int last)
    {
        double buf[MAXD];
        Sizes z;
        z.nv = gP->nv, z.nq = gP->nq, z.nj = gP->nj, z.nf = gP->nframes, z.nc = gP->ncontacts;
        z.n = S.n, z.mi = S.mi, z.rows_dense = S.rows_dense, z.nref = gP->nref, z.ldJ = S.ldJ, z.ldM = S.ldM;
        z.S = &S, z.P = gP;
        const Sm sm = make_sm(buf, z);
        for (int j = 0; j < z.nj; ++j) {
            const int c0 = gP->idx_v[j], cn = gP->jtype[j] == IWBC_J_FREEFLYER ? 6 : 1;
            for (int k = 0; k < cn; ++k)
                sm.dofj[c0 + k] = j;
        }
        int flip = 0;
        // positions [first, last) of the batch / list: the host launches one pass per kernel (below)
        const int total = io.list ? *io.list_count : io.B;
        const int count = last < total ? last : total;
        const int lane = threadIdx.x & 31;
        for (int base = first + blockIdx.x * blockDim.x + (threadIdx.x - lane); base < count; base += gridDim.x * blockDim.x) {
            const int idx = base + lane;
            if (idx < count) {
                const int inst = io.list ? io.list[idx] : idx;
                cycle_instance<1>(gP, sm, z, io, inst, 0, flip);
            }
            __syncwarp();
        }
    }

    // ---- dense QP kernel: the SolverHQPBase-shaped entry point.  H[B][n][n] etc. in global memory.
    template <int G, int IPC>
    __global__ void __launch_bounds__(G* IPC) k_qp_dense(const __grid_constant__ Shape S, QpIO io)
    {
        extern __shared__ double smem[];
        const int grp = threadIdx.x / G;
        const int tid = threadIdx.x - grp * G;
        Sizes z;
        z.nv = 0, z.nq = 0, z.nj = 0, z.nf = 0, z.nc = 0, z.nref = 0, z.rows_dense = 0;
        z.n = io.n;
        z.mi = io.mi;
        z.ldJ = S.ldJ;
        z.ldM = 1;
        z.S = &S;
        z.P = nullptr;
        const Sm sm = make_sm(smem + (size_t)grp * S.total, z);
        const int n = io.n;
        int flip = 0;
        const int ngroups = gridDim.x * IPC;
        for (int inst = blockIdx.x * IPC + grp; inst < io.B; inst += ngroups) {
            const double* H = io.H + (size_t)inst * n * n;
            for (int e = tid; e < n * n; e += G) {
                const int i = e / n, j = e - i * n;
                if (i >= j)
                    sm.J[i + j * z.ldJ] = H[e];
            }
            for (int k = tid; k < n; k += G)
                sm.g[k] = io.g[(size_t)inst * n + k];
            Grp<G>::sync();
            DenseRows rows;
            rows.CE = io.CE + (size_t)inst * io.me * n;
            rows.ce0 = io.ce0 + (size_t)inst * io.me;
            rows.CI = io.CI + (size_t)inst * io.mi * n;
            rows.ci0 = io.ci0 + (size_t)inst * io.mi;
            rows.n = n, rows.me = io.me, rows.mi = io.mi;
            int iters = 0, iq = 0;
            const int status = gi_solve<G, DenseRows>(nullptr, sm, z, &rows, io.me, io.mi, tid, flip, iters, iq);
            Grp<G>::sync();
            for (int k = tid; k < n; k += G)
                io.x[(size_t)inst * n + k] = sm.x[k];
            if (tid == 0) {
                io.status[inst] = status;
                io.iters[inst] = iters;
                unsigned* a = io.active + (size_t)inst * IWBC_ACTIVE_WORDS;
                for (int k = 0; k < IWBC_ACTIVE_WORDS; ++k)
                    a[k] = 0;
                for (int k = io.me; k < iq; ++k) {
                    const int r = sm.act[k];
                    if (r >= 0 && r < 32 * IWBC_ACTIVE_WORDS)
                        a[r >> 5] |= 1u << (r & 31);
                }
            }
            Grp<G>::sync();
        }
    }

} // namespace dev
} // namespace iwbc

using namespace iwbc;

namespace {

#ifdef IWBC_CYC_G
template <int G, int IPC, int MINB = (G <= 32 && G * IPC <= 128 ? 3 : 1)>
int launch_cycle(iwbc_batch* b, const dev::IO& io, cudaStream_t st)
{
    auto kern = dev::k_cycle<G, IPC, MINB>;
    static size_t smem_set[16];
    int rc0 = iwbc_raise_smem_limit((const void*)kern, b->device, b->smem_bytes, smem_set);
    if (rc0)
        return rc0;
    if (b->grid == 0) {
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G * IPC, b->smem_bytes));
        if (per_sm < 1) {
            set_error("cycle kernel does not fit on an SM (shared memory " + std::to_string(b->smem_bytes) + " B)");
            return IWBC_ERR_CUDA;
        }
        b->grid = b->sm_count * per_sm;
    }
    int grid = std::min(b->grid, (io.B + IPC - 1) / IPC);
    dev::IO iop = io;
    if (io.in_list) // a behaviour's bucket: the generic kernel walks it as its instance list
        iop.list = io.in_list, iop.list_count = io.in_count;
    if (b->profiling)
        iop.prof = b->d_prof;
    if (!io.dbg_M && !io.dbg_cycles) {
        iop.work = b->d_slow_count + 2 * b->slot + 1;
        CUDA_TRY(cudaMemsetAsync(iop.work, 0, sizeof(int), st));
    }
    if (b->timing)
        cudaEventRecord(b->ev0, st);
    kern<<<grid, G * IPC, b->smem_bytes, st>>>(b->d_prob, b->shape, iop);
    if (b->timing) {
        cudaEventRecord(b->ev1, st);
        cudaEventSynchronize(b->ev1);
        float ms = 0;
        cudaEventElapsedTime(&ms, b->ev0, b->ev1);
        b->ms_sum += ms;
        b->ms_n++;
    }
    b->launches++;
    CUDA_TRY(cudaGetLastError());
    return IWBC_OK;
}


template <int G, int IPC>
int launch_qp_dense(const dev::Shape& S, const dev::QpIO& io, int device, int sm_count)
{
    auto kern = dev::k_qp_dense<G, IPC>;
    static size_t smem_set[16];
    const size_t smem = (size_t)S.total * 8 * IPC;
    int rc = iwbc_raise_smem_limit((const void*)kern, device, smem, smem_set);
    if (rc)
        return rc;
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G * IPC, smem));
    if (per_sm < 1) {
        set_error("QP kernel does not fit in shared memory");
        return IWBC_ERR_CUDA;
    }
    const int grid = std::min(sm_count * per_sm, (io.B + IPC - 1) / IPC);
    kern<<<grid, G * IPC, smem>>>(S, io);
    CUDA_TRY(cudaGetLastError());
    return IWBC_OK;
}
#endif

#ifdef IWBC_CYC_TPI
int launch_tpi(iwbc_batch* b, const dev::IO& iop, int grid, cudaStream_t st)
{
    // One pass per launch: the warps of a pass walk the (large) kernel in near lock-step and share the instruction cache; in a
    // grid-stride loop they drift apart pass by pass and the cache thrashes (measured, Tiago: 40 M/s in one pass of 65 536 states,
    // 33 / 22 M/s with two / four passes in one launch -- at one CTA per SM the effect is absent).  Equal passes, each at most the
    // resident capacity.
    const int cap = grid * 128, passes = (iop.B + cap - 1) / cap;
    const int per = (((iop.B + passes - 1) / passes) + 127) & ~127;
    for (int first = 0; first < iop.B; first += per) {
        const int last = std::min(iop.B, first + per);
        dev::k_cycle_tpi<IWBC_CYC_TPI, 4><<<(last - first + 127) / 128, 128, 0, st>>>(b->d_prob, b->shape, iop, first, last);
        b->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return IWBC_OK;
}
#endif

} // namespace

#if defined(IWBC_CYC_G) && IWBC_CYC_G == 32
int iwbc_launch_cycle_32_4_3(iwbc_batch* b, const dev::IO& io, cudaStream_t st) { return launch_cycle<32, 4, 3>(b, io, st); }
int iwbc_launch_cycle_32_4_4(iwbc_batch* b, const dev::IO& io, cudaStream_t st) { return launch_cycle<32, 4, 4>(b, io, st); }
int iwbc_launch_qp_dense_32(const dev::Shape& S, const dev::QpIO& io, int device, int sm_count) { return launch_qp_dense<32, 4>(S, io, device, sm_count); }
#elif defined(IWBC_CYC_G) && IWBC_CYC_G == 64
int iwbc_launch_cycle_64(iwbc_batch* b, const dev::IO& io, cudaStream_t st) { return launch_cycle<64, 1>(b, io, st); }
int iwbc_launch_qp_dense_64(const dev::Shape& S, const dev::QpIO& io, int device, int sm_count) { return launch_qp_dense<64, 1>(S, io, device, sm_count); }
#elif defined(IWBC_CYC_G) && IWBC_CYC_G == 96
int iwbc_launch_cycle_96(iwbc_batch* b, const dev::IO& io, cudaStream_t st) { return launch_cycle<96, 1>(b, io, st); }
int iwbc_launch_qp_dense_96(const dev::Shape& S, const dev::QpIO& io, int device, int sm_count) { return launch_qp_dense<96, 1>(S, io, device, sm_count); }
#elif defined(IWBC_CYC_G) && IWBC_CYC_G == 128
int iwbc_launch_cycle_128(iwbc_batch* b, const dev::IO& io, cudaStream_t st) { return launch_cycle<128, 1>(b, io, st); }
int iwbc_launch_qp_dense_128(const dev::Shape& S, const dev::QpIO& io, int device, int sm_count) { return launch_qp_dense<128, 1>(S, io, device, sm_count); }
#elif defined(IWBC_CYC_TPI) && IWBC_CYC_TPI == 1024
int iwbc_launch_cycle_tpi_1024(iwbc_batch* b, const dev::IO& io, int grid, cudaStream_t st) { return launch_tpi(b, io, grid, st); }
#elif defined(IWBC_CYC_TPI) && IWBC_CYC_TPI == 1792
int iwbc_launch_cycle_tpi_1792(iwbc_batch* b, const dev::IO& io, int grid, cudaStream_t st) { return launch_tpi(b, io, grid, st); }
#endif
