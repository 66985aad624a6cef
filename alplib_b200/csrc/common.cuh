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


// ---- weighted-histogram building blocks ---------------------------------------------------------------------------------
// How the bin edges are spaced, detected once per CTA from the staged edges; a spacing guess is always verified against
// the actual edges (np.histogram semantics stay exact), it only replaces the binary search.
struct BinGuess { int mode; double x0, scale; };      // mode 0: none, 1: linear, 2: logarithmic (guess from float log2)

__device__ __forceinline__ BinGuess detect_spacing(const double* s_edges, int n_edges) {
    // (call with all threads of the CTA; s_edges visible)
    const int nb = n_edges - 1;
    const double e0 = s_edges[0], e1 = s_edges[nb];
    bool lin = nb >= 1 && e1 > e0, lg = lin && e0 > 0.0;
    const double w = (e1 - e0) / nb;
    const double l0 = lg ? (double)__log2f((float)e0) : 0.0, lw = lg ? ((double)__log2f((float)e1) - l0) / nb : 1.0;
    for (int i = threadIdx.x; i <= nb; i += blockDim.x) {
        const double e = s_edges[i];
        lin = lin && fabs(e - (e0 + i * w)) <= 0.25 * w;
        lg = lg && e > 0.0 && fabs(((double)__log2f((float)e) - l0) - i * lw) <= 0.25 * lw;
    }
    const int all_lin = __syncthreads_and(lin), all_lg = __syncthreads_and(lg);
    BinGuess g;
    if (all_lin) { g.mode = 1; g.x0 = e0; g.scale = 1.0 / w; }
    else if (all_lg) { g.mode = 2; g.x0 = l0; g.scale = 1.0 / lw; }
    else { g.mode = 0; g.x0 = 0.0; g.scale = 0.0; }
    return g;
}

// np.histogram bin of x (as find_bin), trying the spacing guess first
__device__ __forceinline__ int find_bin_guess(double x, const double* edges, int n_edges, const BinGuess& g) {
    if (g.mode != 0) {
        if (!(x >= edges[0]) || !(x <= edges[n_edges - 1])) return -1;
        const double pos = g.mode == 1 ? (x - g.x0) * g.scale : ((double)__log2f((float)x) - g.x0) * g.scale;
        int b = (int)pos;
        b = b < 0 ? 0 : (b > n_edges - 2 ? n_edges - 2 : b);
        // exact placement: at most one step either way when the guess is good
        if (x < edges[b]) { if (b > 0 && x >= edges[b - 1]) return b - 1; }
        else if (x < edges[b + 1] || b == n_edges - 2) return b;
        else if (b + 2 <= n_edges - 1 && (x < edges[b + 2] || b + 2 == n_edges - 1)) return b + 1;
    }
    return find_bin(x, edges, n_edges);
}

// Branch-free bin lookup for the fused histogram of the row kernel.  The spacing guess is evaluated in float (one FFMA on x
// or on log2 x, floor), the bin's edge pair {edges[b], edges[b+1]} comes with ONE 16-byte shared load from `pairs`
// (pairs[nb-1].y = nextafter(edges[nb], +inf): np.histogram's closed last bin), and two compares confirm the bin.  Returns
// the bin, -1 when x lies outside [edges[0], edges[nb]] (or is NaN), or -2 when x is in range but the guess missed (float
// rounding next to an edge, irregular edges): the caller redoes those rare lanes with find_bin().
struct BinGuessF { int mode; float off, scale; double lo, hi; };

__device__ __forceinline__ BinGuessF to_float_guess(const BinGuess& g, const double* s_edges, int n_edges) {
    BinGuessF f;
    f.mode = g.mode;
    f.scale = (float)g.scale;
    f.off = (float)(-g.x0 * g.scale);
    f.lo = s_edges[0];
    f.hi = s_edges[n_edges - 1];
    return f;
}

__device__ __forceinline__ int find_bin_pairs(double x, const double2* pairs, int nb, const BinGuessF& g) {
    float v = (float)x;
    if (g.mode == 2) {                    // (a real, warp-uniform branch: linear edges do not pay for the logarithm)
#if defined(__CUDA_ARCH__)
        asm volatile("");
#endif
        v = __log2f(v);
    }
    int b = __float2int_rd(fmaf(v, g.scale, g.off));
    b = max(0, min(b, nb - 1));
    const double2 e = pairs[b];
    const bool hit = (x >= e.x) & (x < e.y);
    const bool inside = (x >= g.lo) & (x <= g.hi);
    return hit ? b : (inside ? -2 : -1);
}

// Add w to bins[b] for the valid lanes of a fully converged warp, without atomics: `bins` is private to this warp.  Lanes
// that hit the same bin are combined first (match.any); the leader of each group does a plain read-modify-write.  Two
// schedules: few distinct bins -> one masked butterfly sum per bin; many distinct bins -> the leaders walk their peers.
__device__ __forceinline__ void warp_hist_add(double* bins, int b, double w, bool valid) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(full, valid ? b : -1 - lane);       // invalid lanes: unique negative keys
    const bool leader = valid && (lane == __ffs(peers) - 1);
    const unsigned leaders = __ballot_sync(full, leader);
    if (leaders == 0) return;
    const int n_groups = __popc(leaders);
    const int max_group = __reduce_max_sync(full, valid ? __popc(peers) : 0);
    if (n_groups <= max_group) {
        unsigned todo = leaders;
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            const unsigned grp = __shfl_sync(full, peers, l);
            double v = ((grp >> lane) & 1u) ? w : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
            if (lane == l) bins[b] += v;
        }
    } else {
        double acc = w;
        unsigned rem = peers & ~(1u << lane);
        for (int r = 1; r < max_group; ++r) {
            const bool have = rem != 0;
            const int src = have ? __ffs(rem) - 1 : lane;
            rem &= rem - 1;
            const double v = __shfl_sync(full, w, src);
            if (leader && have) acc += v;
        }
        if (leader) bins[b] += acc;
    }
}

// Stage `nrows` records of `NP` doubles starting at row r0 of a row-major table into shared memory.
template <int NP>
__device__ __forceinline__ void stage_rows(double* dst, const double* __restrict__ table, uint64_t r0, int nrows) {
    const double* src = table + r0 * NP;
    for (int i = threadIdx.x; i < nrows * NP; i += blockDim.x) dst[i] = __ldg(src + i);
}
#endif

}  // namespace alpk
