// fast_tu.cuh -- one translation unit per compile-time (nv, me) shape of the k_fast / k_slow pair, so that the shapes
// compile in parallel.  The kernels take the launch's shape tables as __grid_constant__ parameters and the problem / plan from the
// handle's own global-memory copies: no per-unit __constant__ state, nothing to refresh when another handle or mask launches.
// Usage: #define IWBC_TU_NV / IWBC_TU_ME / IWBC_TU_ENTRY, then #include "fast_tu.cuh".
#include <algorithm>

#include "batch_state.hpp"

using namespace iwbc;

namespace {

size_t g_fast_smem_set[16], g_slow_smem_set[16];

constexpr int FT_ = IWBC_FAST_THREADS;

template <int NV_, int ME_>
int launch_fast_shape(iwbc_batch* b, const dev::IO& io, cudaStream_t st)
{
    dev::IO iop = io;
    auto fk = dev::k_fast<FT_, NV_, ME_>;
    auto sk = dev::k_slow<FT_, NV_, ME_>;
    int rc = iwbc_raise_smem_limit((const void*)fk, b->device, b->fast_smem, g_fast_smem_set);
    if (rc == IWBC_OK)
        rc = iwbc_raise_smem_limit((const void*)sk, b->device, b->slow_smem, g_slow_smem_set);
    if (rc)
        return rc;
    if (b->fast_grid == 0) {
        int per_sm = 0, per_sm_slow = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fk, FT_, b->fast_smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_slow, sk, FT_, b->slow_smem));
        if (per_sm < 1 || per_sm_slow < 1) {
            set_error("fast / active-set kernels do not fit on an SM");
            return IWBC_ERR_CUDA;
        }
        b->fast_grid = b->sm_count * per_sm;
        b->slow_grid = b->sm_count * per_sm_slow;
    }
    iop.slow_list = b->d_slow_list + b->list_off;
    iop.slow_count = b->d_slow_count + 2 * b->slot; // [2 slot]: length of the slow list, [2 slot + 1]: k_slow's work counter
    if (b->d_scratch == nullptr || b->scratch_img < b->fshape.total) {
        // (re)allocate the hand-over scratch (grow only): at most 65536 images (70 kB each for Talos double support)
        if (b->d_scratch) {
            CUDA_TRY(cudaDeviceSynchronize());
            cudaFree(b->d_scratch);
            b->d_scratch = nullptr;
        }
        b->scratch_cap = std::min(b->max_batch, 65536);
        b->scratch_img = b->fshape.total + 64;
        CUDA_TRY(cudaMalloc((void**)&b->d_scratch, (size_t)b->scratch_cap * b->scratch_img * sizeof(double)));
    }
    {
        const int off = b->scratch_len ? b->scratch_off : 0, len = b->scratch_len ? b->scratch_len : b->scratch_cap;
        iop.scratch = b->d_scratch + (size_t)off * b->scratch_img; // (images are packed with the current plan's size inside this range)
        iop.slow_code = b->d_slow_code + off;
        iop.scratch_cap = len;
    }
    if (b->profiling)
        iop.prof = b->d_prof;
    CUDA_TRY(cudaMemsetAsync(b->d_slow_count + 2 * b->slot, 0, 2 * sizeof(int), st));
    if (b->timing)
        cudaEventRecord(b->ev0, st);
    fk<<<std::min(b->fast_grid, io.B), FT_, b->fast_smem, st>>>(b->d_prob, b->d_fplan, b->shape, b->fshape, iop);
    b->launches++;
    iop.list = b->d_slow_list + b->list_off;
    iop.list_count = b->d_slow_count + 2 * b->slot;
    iop.work = b->d_slow_count + 2 * b->slot + 1;
    sk<<<std::min(b->slow_grid, io.B), FT_, b->slow_smem, st>>>(b->d_prob, b->d_fplan, b->shape, b->fshape, iop);
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

} // namespace

int IWBC_TU_ENTRY(iwbc_batch* b, const dev::IO& io, cudaStream_t st) { return launch_fast_shape<IWBC_TU_NV, IWBC_TU_ME>(b, io, st); }
