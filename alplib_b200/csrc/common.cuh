// Launch plumbing shared by the alpk kernels: error reporting, grid sizing, block reductions, row-table staging.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "alp_math.cuh"

#define ALPK_EXPORT extern "C" __attribute__((visibility("default")))

namespace alpk {

constexpr int kThreads = 256;            // threads per CTA for the per-sample kernels
constexpr int kMaxPartials = 4096;       // caller-provided scratch: >= kMaxPartials doubles per reduced quantity
constexpr int kStageRows = 48;           // per-row parameter records staged in shared memory per tile

void set_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

// Persistent-style grid: a multiple of the SM count (148 on B200), capped by the amount of work.
inline int grid_for(uint64_t n_items, int items_per_block, int blocks_per_sm) {
    uint64_t need = (n_items + (uint64_t)items_per_block - 1) / (uint64_t)items_per_block;
    uint64_t cap = (uint64_t)sm_count() * (uint64_t)blocks_per_sm;
    if (cap > (uint64_t)kMaxPartials) cap = kMaxPartials;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

#if defined(__CUDACC__)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA; result valid in thread 0.  `red` is a shared array of >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (wid == 0) v = warp_sum(v);
    return v;
}

// np.histogram bin lookup: bin k = [edges[k], edges[k+1]), last bin closed; -1 outside (also NaN)
__device__ __forceinline__ int find_bin(double x, const double* edges, int n_edges) {
    if (!(x >= edges[0]) || !(x <= edges[n_edges - 1])) return -1;    // also drops NaN
    if (x == edges[n_edges - 1]) return n_edges - 2;
    int lo = 0, hi = n_edges - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= edges[mid]) lo = mid; else hi = mid;
    }
    return lo;
}


// Stage `nrows` records of `NP` doubles starting at row r0 of a row-major table into shared memory.
template <int NP>
__device__ __forceinline__ void stage_rows(double* dst, const double* __restrict__ table, uint64_t r0, int nrows) {
    const double* src = table + r0 * NP;
    for (int i = threadIdx.x; i < nrows * NP; i += blockDim.x) dst[i] = __ldg(src + i);
}
#endif

}  // namespace alpk
