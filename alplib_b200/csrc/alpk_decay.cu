// libalpk: a -> gamma gamma two-body decay MC (lab-frame 4-vectors) and the weighted histogram reduction.
#include <string.h>

#include "../../include/alpk.h"
#include "common.cuh"

using namespace alp;

namespace alpk {

// ---------------------------------------------------------------------------------------------------------------
// per-parent constants (generators.py:97,117-121; cross_section_mc.py:510-525,676-685)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
decay_parent_rows_kernel(const double* __restrict__ energy, const float* __restrict__ w32, uint64_t n, double ma2,
                         EventConst ev, float n_samples32, double n_samples64, double* __restrict__ table) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float w = __ldg(w32 + i);
        // event_weight = days*S_PER_DAY * decay_axion_weight / n_samples   (generators.py:97; float32 unless the
        // exposure is an np.float64)
        const double ew = ev.days_f64 ? (ev.days_sec * (double)w) / n_samples64
                                      : (double)((ev.days_sec32 * w) / n_samples32);
        double row[PR_NP];
        decay_parent_params(__ldg(energy + i), ma2, ew, row);
#pragma unroll
        for (int k = 0; k < PR_NP; ++k) table[i * PR_NP + k] = row[k];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-event kernel: 72 B written per event (2 x 4 x f64 + f64 weight), SoA planes, 16-byte stores (event pairs)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDUnroll = 2;
constexpr int kDTile = kThreads * 2 * kDUnroll;

struct DecayArgs {
    const double* table; uint64_t n_parents; uint32_t n_samples; const uint32_t* row_ids;
    const double* uniforms; uint64_t seed;
    double* p1; double* p2; double* w;
};

template <bool INJECT, bool FAST>
__global__ void __launch_bounds__(kThreads)
decay2body_kernel(const DecayArgs a) {
    __shared__ double s_rows[kStageRows * PR_NP];
    const uint64_t n = a.n_parents * (uint64_t)a.n_samples;
    const uint64_t n_tiles = (n + kDTile - 1) / kDTile;
    const bool planes_aligned = (n & 1ull) == 0;     // odd N shifts the odd planes off 16-byte alignment

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t base = tile * kDTile;
        const uint64_t last = (base + kDTile <= n ? base + kDTile : n) - 1;
        const uint64_t r0 = base / a.n_samples;
        const uint64_t r1 = last / a.n_samples;
        const int nrows = (int)(r1 - r0 + 1);
        const double* rows;
        __syncthreads();
        if (nrows <= kStageRows) {
            stage_rows<PR_NP>(s_rows, a.table, r0, nrows);
            rows = s_rows;
        } else {
            rows = a.table + r0 * PR_NP;
        }
        __syncthreads();
        const uint64_t row0_first = r0 * (uint64_t)a.n_samples;

#pragma unroll 1
        for (int k = 0; k < kDUnroll; ++k) {
            const uint64_t i0 = base + 2 * ((uint64_t)k * kThreads + threadIdx.x);
            if (i0 >= n) break;
            const bool have2 = (i0 + 1 < n);
            double q1[2][4], q2[2][4], wv[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint32_t off = (uint32_t)(i0 + ((j == 1 && !have2) ? 0 : j) - row0_first);
                const uint32_t rr = off / a.n_samples;
                const uint32_t s = off - rr * a.n_samples;
                const double* p = rows + (size_t)rr * PR_NP;
                const uint64_t r = r0 + rr;
                double u1, u2;
                if (INJECT) {
                    const double* ub = a.uniforms + (r * 2) * (uint64_t)a.n_samples;
                    u1 = __ldg(ub + s);
                    u2 = __ldg(ub + a.n_samples + s);
                } else {
                    const uint32_t rid = a.row_ids ? __ldg(a.row_ids + r) : (uint32_t)r;
                    uniform_2(a.seed, kTagDecay2, rid, s, u1, u2);
                }
                if (FAST) decay_event_fast(u1, u2, p, q1[j], q2[j]);
                else decay_event(u1, u2, p, q1[j], q2[j]);
                wv[j] = p[PR_W];
            }
            if (have2 && planes_aligned) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    *reinterpret_cast<double2*>(a.p1 + (uint64_t)c * n + i0) = make_double2(q1[0][c], q1[1][c]);
                    *reinterpret_cast<double2*>(a.p2 + (uint64_t)c * n + i0) = make_double2(q2[0][c], q2[1][c]);
                }
                *reinterpret_cast<double2*>(a.w + i0) = make_double2(wv[0], wv[1]);
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    a.p1[(uint64_t)c * n + i0] = q1[0][c];
                    a.p2[(uint64_t)c * n + i0] = q2[0][c];
                }
                a.w[i0] = wv[0];
                if (have2) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        a.p1[(uint64_t)c * n + i0 + 1] = q1[1][c];
                        a.p2[(uint64_t)c * n + i0 + 1] = q2[1][c];
                    }
                    a.w[i0 + 1] = wv[1];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weighted histogram, np.histogram(x, bins=edges, weights=w): bin k = [edges[k], edges[k+1]), last bin closed
// ---------------------------------------------------------------------------------------------------------------
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

template <bool SMEM>
__global__ void __launch_bounds__(kThreads)
hist_kernel(const double* __restrict__ x, const double* __restrict__ w64, const float* __restrict__ w32, uint64_t n,
            const double* __restrict__ edges, int n_edges, double* __restrict__ hist) {
    extern __shared__ double sm[];
    const int n_bins = n_edges - 1;
    const double* ed = edges;
    double* h = hist;
    if (SMEM) {
        double* s_edges = sm;
        double* s_hist = sm + n_edges;
        for (int i = threadIdx.x; i < n_edges; i += blockDim.x) s_edges[i] = __ldg(edges + i);
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x) s_hist[i] = 0.0;
        __syncthreads();
        ed = s_edges;
        h = s_hist;
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int b = find_bin(__ldg(x + i), ed, n_edges);
        if (b >= 0) {
            const double w = w64 ? __ldg(w64 + i) : (w32 ? (double)__ldg(w32 + i) : 1.0);
            atomicAdd(h + b, w);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x) {
            const double v = h[i];
            if (v != 0.0) atomicAdd(hist + i, v);
        }
    }
}

}  // namespace alpk

using namespace alpk;

ALPK_EXPORT int alpk_decay_parent_rows(const double* energy, const float* decay_w32, uint64_t n_parents, double ma2,
                                       const alpk_event_t* h_ev, uint32_t n_samples, double* row_table, void* stream) {
    static_assert(PR_NP == ALPK_PARENT_NP, "header/table layout mismatch");
    if (!h_ev || !row_table) { set_error("alpk_decay_parent_rows: null argument"); return 2; }
    if (n_parents == 0) return 0;
    EventConst ev;
    ev.days_sec = h_ev->days_sec; ev.days_sec32 = h_ev->days_sec32; ev.days_f64 = h_ev->days_f64; ev.threshold = h_ev->threshold;
    decay_parent_rows_kernel<<<grid_for(n_parents, kThreads, 4), kThreads, 0, (cudaStream_t)stream>>>(
        energy, decay_w32, n_parents, ma2, ev, (float)n_samples, (double)n_samples, row_table);
    return check_launch("alpk_decay_parent_rows");
}

ALPK_EXPORT int alpk_decay2body(const double* row_table, uint64_t n_parents, uint32_t n_samples, const uint32_t* row_ids,
                                const double* uniforms, uint64_t seed, int fast, double* p1, double* p2, double* w, void* stream) {
    if (!row_table || !p1 || !p2 || !w) { set_error("alpk_decay2body: null argument"); return 2; }
    if (n_samples == 0 || n_samples > 0x7fffffffu - kDTile) { set_error("alpk_decay2body: n_samples out of range"); return 2; }
    const uint64_t n = n_parents * (uint64_t)n_samples;
    if (n == 0) return 0;
    DecayArgs a;
    a.table = row_table; a.n_parents = n_parents; a.n_samples = n_samples; a.row_ids = row_ids;
    a.uniforms = uniforms; a.seed = seed; a.p1 = p1; a.p2 = p2; a.w = w;
    const int grid = grid_for(n, kDTile, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (uniforms) { if (fast) decay2body_kernel<true, true><<<grid, kThreads, 0, st>>>(a); else decay2body_kernel<true, false><<<grid, kThreads, 0, st>>>(a); }
    else { if (fast) decay2body_kernel<false, true><<<grid, kThreads, 0, st>>>(a); else decay2body_kernel<false, false><<<grid, kThreads, 0, st>>>(a); }
    return check_launch("alpk_decay2body");
}

ALPK_EXPORT int alpk_hist(const double* x, const double* w64, const float* w32, uint64_t n, const double* edges,
                          int n_edges, double* hist, void* stream) {
    if (n_edges < 2) { set_error("alpk_hist: need at least 2 edges"); return 2; }
    if (!x || !edges || !hist) { set_error("alpk_hist: null argument"); return 2; }
    if (n == 0) return 0;
    const size_t smem = (size_t)(2 * n_edges - 1) * sizeof(double);
    const int grid = grid_for(n, kThreads * 8, 4);
    if (smem <= 48 * 1024) hist_kernel<true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(x, w64, w32, n, edges, n_edges, hist);
    else hist_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, w64, w32, n, edges, n_edges, hist);
    return check_launch("alpk_hist");
}
