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
// optional FP32 mode of the decay MC (north star: 1e-5 relative): per-parent constants stay FP64 (they contain the
// cancellations), per-event arithmetic, RNG mantissas (24 bit) and outputs are float -> 36 B/event instead of 72.
// thread = 4 consecutive events: two Philox calls, 16-byte float4 stores per plane.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
decay2body_f32_kernel(const DecayArgs a, float* __restrict__ p1, float* __restrict__ p2, float* __restrict__ w) {
    const uint64_t n = a.n_parents * (uint64_t)a.n_samples;
    const uint64_t n_quads = (n + 3) / 4;
    const bool aligned = (n & 3ull) == 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t qd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; qd < n_quads; qd += stride) {
        const uint64_t i0 = qd * 4;
        float q1[4][4], q2[4][4], wv[4];
        // (row, sample) of the first event by one 64-bit division, the next three by increment
        uint64_t r = i0 / a.n_samples;
        uint32_t s = (uint32_t)(i0 - r * a.n_samples);
        Philox4 x = {0u, 0u, 0u, 0u};
        uint32_t have_ctr = 0xffffffffu; uint64_t have_row = ~0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j > 0 && i0 + j < n) { if (++s == a.n_samples) { s = 0; ++r; } }
            const double* p = a.table + r * PR_NP;
            // event s uses Philox counter s>>1 and word pair s&1 (24-bit mantissas): neighbours share the call
            if ((s >> 1) != have_ctr || r != have_row) {
                const uint32_t rid = a.row_ids ? __ldg(a.row_ids + r) : (uint32_t)r;
                x = philox4x32_10(s >> 1, 0u, rid, kTagDecay2 + 16u, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                have_ctr = s >> 1; have_row = r;
            }
            const float u1 = (float)(((s & 1u) ? x.z : x.x) >> 8) * 5.9604644775390625e-8f;
            const float u2 = (float)(((s & 1u) ? x.w : x.y) >> 8) * 5.9604644775390625e-8f;
            float sphi, cphi;
            sincospif(2.0f * u1, &sphi, &cphi);
            const float cth = 1.0f - 2.0f * u2;
            const float sth = 2.0f * sqrtf(u2 * (1.0f - u2));
            const float pcm = (float)__ldg(p + PR_PCM), ecm = (float)__ldg(p + PR_ECM), gam = (float)__ldg(p + PR_GAM);
            const float m03 = (float)__ldg(p + PR_M03), m33 = (float)__ldg(p + PR_M33);
            const float px = pcm * cphi * sth, py = pcm * sphi * sth, pz = pcm * cth;
            const float ge = gam * ecm, me = m03 * ecm, mp = m03 * pz, m3p = m33 * pz;
            q1[j][0] = ge + mp; q1[j][1] = px;  q1[j][2] = py;  q1[j][3] = me + m3p;
            q2[j][0] = ge - mp; q2[j][1] = -px; q2[j][2] = -py; q2[j][3] = me - m3p;
            wv[j] = (float)__ldg(p + PR_W);
        }
        if (aligned) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                *reinterpret_cast<float4*>(p1 + (uint64_t)c * n + i0) = make_float4(q1[0][c], q1[1][c], q1[2][c], q1[3][c]);
                *reinterpret_cast<float4*>(p2 + (uint64_t)c * n + i0) = make_float4(q2[0][c], q2[1][c], q2[2][c], q2[3][c]);
            }
            *reinterpret_cast<float4*>(w + i0) = make_float4(wv[0], wv[1], wv[2], wv[3]);
        } else {
            for (int j = 0; j < 4 && i0 + j < n; ++j) {
                for (int c = 0; c < 4; ++c) { p1[(uint64_t)c * n + i0 + j] = q1[j][c]; p2[(uint64_t)c * n + i0 + j] = q2[j][c]; }
                w[i0 + j] = wv[j];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// general 2-body decay (arbitrary parent direction, massive daughters): one thread per event, dense 4x4 boost
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
decay_general_rows_kernel(const double* __restrict__ p4, const double* __restrict__ w, uint64_t n, double m1, double m2,
                          double* __restrict__ table) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double row[PG_NP];
        decay_parent_params_general(__ldg(p4 + 4 * i), __ldg(p4 + 4 * i + 1), __ldg(p4 + 4 * i + 2), __ldg(p4 + 4 * i + 3),
                                    m1, m2, w ? __ldg(w + i) : 1.0, row);
        for (int k = 0; k < PG_NP; ++k) table[i * PG_NP + k] = row[k];
    }
}

__global__ void __launch_bounds__(kThreads)
decay_general_kernel(const double* __restrict__ table, uint64_t n_parents, uint32_t n_samples,
                     const uint32_t* __restrict__ row_ids, const double* __restrict__ uniforms, uint64_t seed,
                     double* __restrict__ p1, double* __restrict__ p2, double* __restrict__ wout,
                     double* __restrict__ theta1) {
    const uint64_t n = n_parents * (uint64_t)n_samples;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t r = i / n_samples;
        const uint32_t s = (uint32_t)(i - r * n_samples);
        double row[PG_NP];
        for (int k = 0; k < PG_NP; ++k) row[k] = __ldg(table + r * PG_NP + k);
        double u1, u2;
        if (uniforms) {
            u1 = __ldg(uniforms + (r * 2) * n_samples + s);
            u2 = __ldg(uniforms + (r * 2 + 1) * n_samples + s);
        } else {
            uniform_2(seed, kTagDecay2, row_ids ? __ldg(row_ids + r) : (uint32_t)r, s, u1, u2);
        }
        double q1[4], q2[4];
        const double th = decay_event_general(u1, u2, row, q1, q2);
        for (int c = 0; c < 4; ++c) {
            p1[(uint64_t)c * n + i] = q1[c];
            if (p2) p2[(uint64_t)c * n + i] = q2[c];
        }
        if (wout) wout[i] = row[PG_W];
        if (theta1) theta1[i] = th;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// weighted histogram, np.histogram(x, bins=edges, weights=w): bin k = [edges[k], edges[k+1]), last bin closed
// ---------------------------------------------------------------------------------------------------------------
// WARP_PRIVATE: every warp owns a copy of the bins in shared memory and adds to it without atomics (warp_hist_add);
// otherwise (too many bins for 8 copies) one copy per CTA with shared atomics, or global atomics when even that does not
// fit.  Four samples per trip with the loads issued first.
template <int MODE>      // 0: global atomics, 1: per-CTA shared bins, 2: per-warp shared bins
__global__ void __launch_bounds__(kThreads)
hist_kernel(const double* __restrict__ x, const double* __restrict__ w64, const float* __restrict__ w32, uint64_t n,
            const double* __restrict__ edges, int n_edges, double* __restrict__ hist) {
    extern __shared__ double sm[];
    const int n_bins = n_edges - 1;
    const int copies = MODE == 2 ? kThreads / 32 : 1;
    const double* ed = edges;
    double* h = hist;
    BinGuess guess{0, 0.0, 0.0};
    if (MODE >= 1) {
        double* s_edges = sm;
        double* s_hist = sm + n_edges;
        for (int i = threadIdx.x; i < n_edges; i += blockDim.x) s_edges[i] = __ldg(edges + i);
        for (int i = threadIdx.x; i < n_bins * copies; i += blockDim.x) s_hist[i] = 0.0;
        __syncthreads();
        guess = detect_spacing(s_edges, n_edges);
        ed = s_edges;
        h = s_hist + (MODE == 2 ? (threadIdx.x >> 5) * n_bins : 0);
    }
    constexpr int U = 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // whole warps iterate together (warp_hist_add needs every lane): trip count from the warp's first lane
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
    for (uint64_t base = first; base < n; base += U * stride) {
        double xv[U], wv[U];
        bool ok[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const uint64_t i = base + (threadIdx.x & 31) + q * stride;
            ok[q] = i < n;
            xv[q] = ok[q] ? __ldg(x + i) : 0.0;
            wv[q] = !ok[q] ? 0.0 : (w64 ? __ldg(w64 + i) : (w32 ? (double)__ldg(w32 + i) : 1.0));
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            if (base + q * stride >= n) break;                       // warp-uniform
            const int b = ok[q] ? find_bin_guess(xv[q], ed, n_edges, guess) : -1;
            if (MODE == 2) warp_hist_add(h, b, wv[q], b >= 0);
            else if (b >= 0) atomicAdd(h + b, wv[q]);
        }
    }
    if (MODE >= 1) {
        __syncthreads();
        double* s_hist = sm + n_edges;
        for (int i = threadIdx.x; i < n_bins; i += blockDim.x) {
            double v = 0.0;
            for (int c = 0; c < copies; ++c) v += s_hist[c * n_bins + i];
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
    if (!energy || !decay_w32) { set_error("alpk_decay_parent_rows: null input"); return 2; }
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
    // the kernel stores pairs of events as 16-byte vectors into every plane (odd n is handled in the kernel: scalar stores)
    if (((uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)w) & 15u) { set_error("alpk_decay2body: p1, p2, w must be 16-byte aligned"); return 2; }
    DecayArgs a;
    a.table = row_table; a.n_parents = n_parents; a.n_samples = n_samples; a.row_ids = row_ids;
    a.uniforms = uniforms; a.seed = seed; a.p1 = p1; a.p2 = p2; a.w = w;
    const int grid = grid_for(n, kDTile, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (uniforms) { if (fast) decay2body_kernel<true, true><<<grid, kThreads, 0, st>>>(a); else decay2body_kernel<true, false><<<grid, kThreads, 0, st>>>(a); }
    else { if (fast) decay2body_kernel<false, true><<<grid, kThreads, 0, st>>>(a); else decay2body_kernel<false, false><<<grid, kThreads, 0, st>>>(a); }
    return check_launch("alpk_decay2body");
}

ALPK_EXPORT int alpk_decay2body_f32(const double* row_table, uint64_t n_parents, uint32_t n_samples, const uint32_t* row_ids,
                                    uint64_t seed, float* p1, float* p2, float* w, void* stream) {
    if (!row_table || !p1 || !p2 || !w) { set_error("alpk_decay2body_f32: null argument"); return 2; }
    if (n_samples == 0) { set_error("alpk_decay2body_f32: n_samples out of range"); return 2; }
    const uint64_t n = n_parents * (uint64_t)n_samples;
    if (n == 0) return 0;
    if (((uintptr_t)p1 | (uintptr_t)p2 | (uintptr_t)w) & 15u) { set_error("alpk_decay2body_f32: p1, p2, w must be 16-byte aligned"); return 2; }
    DecayArgs a;
    a.table = row_table; a.n_parents = n_parents; a.n_samples = n_samples; a.row_ids = row_ids;
    a.uniforms = nullptr; a.seed = seed; a.p1 = nullptr; a.p2 = nullptr; a.w = nullptr;
    decay2body_f32_kernel<<<grid_for((n + 3) / 4, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(a, p1, p2, w);
    return check_launch("alpk_decay2body_f32");
}

ALPK_EXPORT int alpk_decay2body_general(const double* parents_p4, const double* parent_w, uint64_t n_parents, double m1,
                                        double m2, uint32_t n_samples, const uint32_t* row_ids, const double* uniforms,
                                        uint64_t seed, double* row_table, double* p1, double* p2, double* w,
                                        double* theta1, void* stream) {
    static_assert(PG_NP == ALPK_PARENT_GENERAL_NP, "header/table layout mismatch");
    if (!parents_p4 || !row_table || !p1) { set_error("alpk_decay2body_general: null argument"); return 2; }
    if (n_parents == 0 || n_samples == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    decay_general_rows_kernel<<<grid_for(n_parents, kThreads, 4), kThreads, 0, st>>>(parents_p4, parent_w, n_parents, m1, m2, row_table);
    if (check_launch("alpk_decay2body_general(rows)")) return 1;
    decay_general_kernel<<<grid_for(n_parents * (uint64_t)n_samples, kThreads, 8), kThreads, 0, st>>>(
        row_table, n_parents, n_samples, row_ids, uniforms, seed, p1, p2, w, theta1);
    return check_launch("alpk_decay2body_general");
}

ALPK_EXPORT int alpk_hist(const double* x, const double* w64, const float* w32, uint64_t n, const double* edges,
                          int n_edges, double* hist, void* stream) {
    if (n_edges < 2) { set_error("alpk_hist: need at least 2 edges"); return 2; }
    if (!x || !edges || !hist) { set_error("alpk_hist: null argument"); return 2; }
    if (n == 0) return 0;
    const size_t smem1 = (size_t)(2 * n_edges - 1) * sizeof(double);
    const size_t smem8 = (size_t)(n_edges + (kThreads / 32) * (n_edges - 1)) * sizeof(double);
    const int grid = grid_for(n, kThreads * 8, 4);
    if (smem8 <= 48 * 1024) hist_kernel<2><<<grid, kThreads, smem8, (cudaStream_t)stream>>>(x, w64, w32, n, edges, n_edges, hist);
    else if (smem1 <= 48 * 1024) hist_kernel<1><<<grid, kThreads, smem1, (cudaStream_t)stream>>>(x, w64, w32, n, edges, n_edges, hist);
    else hist_kernel<0><<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, w64, w32, n, edges, n_edges, hist);
    return check_launch("alpk_hist");
}

// ---------------------------------------------------------------------------------------------------------------
// propagate_iso_vol_int (fluxes.py:67-89): per sample, Monte-Carlo integral over detector-volume points of
//   f(l) = (1/(hbar c v tau)) exp(-l/(hbar c v tau)) / (4 pi l^2)        (materials.py:98-124 DetectorGeometry.integrate)
// thread = sample; the geometry points (host-prepared: -l/METER_BY_MEV, 4 pi l^2, jacobian) are walked from shared
// memory in tiles (broadcast reads); blockIdx.y splits the points, partial sums are reduced in fixed order.
// ---------------------------------------------------------------------------------------------------------------
namespace alpk {

constexpr int kVolTile = 1024;

template <bool FAST>
__global__ void __launch_bounds__(kThreads)
vol_int_kernel(const double* __restrict__ energy, uint64_t n, const double* __restrict__ nxm,
               const double* __restrict__ bden, const double* __restrict__ jac, uint32_t n_points, uint32_t chunk,
               double ma, double ma2, double width, double inv_mbm, double* __restrict__ partials) {
    __shared__ double s_x[kVolTile], s_b[kVolTile], s_j[kVolTile];
    __shared__ double s_tab[alp::kExpTabSize];                     // exp_fast table
    __shared__ unsigned long long s_xmax_bits;       // fast mode: largest |x| of the staged tile (ordered bits)
    if (FAST) for (int i = threadIdx.x; i < alp::kExpTabSize; i += blockDim.x) s_tab[i] = kExpTabDev[i];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t k0 = blockIdx.y * chunk;
    const uint32_t k1 = (k0 + chunk < n_points) ? k0 + chunk : n_points;
    double va = 1.0, tau = 1.0, A = 0.0, c = 0.0;
    const bool live = i < n;
    if (live) {
        const double ea = __ldg(energy + i);
        const double pa = sqrt(ea * ea - ma2);                               // fluxes.py:72
        va = pa / ea;                                                        // :73
        const double boost = ea / ma;                                        // :74
        tau = width > 0.0 ? boost / width : INFINITY;                        // :75
        A = (inv_mbm / va) / tau;                                            // 1/METER_BY_MEV/v/tau   :84
        c = width > 0.0 ? (ma * width) / pa : 0.0;                           // fast: 1/(v tau)
    }
    double acc = 0.0;
    for (uint32_t t0 = k0; t0 < k1; t0 += kVolTile) {
        const uint32_t cnt = (t0 + kVolTile <= k1) ? kVolTile : k1 - t0;
        __syncthreads();
        if (FAST && threadIdx.x == 0) s_xmax_bits = 0ull;
        __syncthreads();
        unsigned long long my_max = 0ull;
        for (uint32_t k = threadIdx.x; k < cnt; k += blockDim.x) {
            const double x = __ldg(nxm + t0 + k), b = __ldg(bden + t0 + k), j = __ldg(jac + t0 + k);
            s_x[k] = x; s_b[k] = b; s_j[k] = j;
            if (FAST) {
                s_b[k] = j / b;                      // fast mode: the per-point factor jacobian / (4 pi l^2), one division per point
                const double ax = fabs(x);           // (NaN -> +inf: no straight-line loop for this tile)
                const unsigned long long xb = (ax == ax) ? (unsigned long long)__double_as_longlong(ax) : 0x7ff0000000000000ull;
                my_max = xb > my_max ? xb : my_max;
            }
        }
        if (FAST && my_max) atomicMax(&s_xmax_bits, my_max);
        __syncthreads();
        if (live) {
            if (FAST && c * __longlong_as_double((long long)s_xmax_bits) <= 708.0) {
                // every exponent of the tile lies in [-708, 0]: straight-line body, four points in flight
#pragma unroll 4
                for (uint32_t k = 0; k < cnt; ++k) acc += exp_fast_core(s_x[k] * c, s_tab) * s_b[k];     // (x A after the walk)
            } else if (FAST) {
#pragma unroll 2
                for (uint32_t k = 0; k < cnt; ++k) acc += exp_fast(s_x[k] * c) * s_b[k];
            } else {
#pragma unroll 4
                for (uint32_t k = 0; k < cnt; ++k) {
                    const double e = exp((s_x[k] / va) / tau);                                // :85
                    acc += s_j[k] * ((A * e) / s_b[k]);
                }
            }
        }
    }
    if (live) partials[(uint64_t)blockIdx.y * n + i] = FAST ? A * acc : acc;       // fast mode: the sample's 1/(hbar c v tau) once
}

template <bool FAST>
__global__ void __launch_bounds__(kThreads)
vol_int_finish_kernel(const double* __restrict__ energy, const double* __restrict__ weight, uint64_t n,
                      const double* __restrict__ partials, uint32_t n_chunks, double vol_over_n, PropConst c,
                      float* __restrict__ d32, float* __restrict__ s32, double* __restrict__ d64o, double* __restrict__ s64o) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double sum = 0.0;
        for (uint32_t q = 0; q < n_chunks; ++q) sum += partials[(uint64_t)q * n + i];
        const double vol = vol_over_n * sum;                      // V * np.sum(...) / n_samples (materials.py:103,110,117)
        double d_unused, s64;
        if (FAST) propagate_sample_fast(__ldg(energy + i), __ldg(weight + i), c, d_unused, s64);
        else propagate_sample(__ldg(energy + i), __ldg(weight + i), c, d_unused, s64);   // scatter weight, fluxes.py:77-78,89
        const double d64 = (c.rescale * __ldg(weight + i)) * vol;                        // :88
        d32[i] = (float)d64;
        s32[i] = (float)s64;
        if (d64o) d64o[i] = d64;
        if (s64o) s64o[i] = s64;
    }
}

}  // namespace alpk

ALPK_EXPORT uint32_t alpk_vol_int_chunks(uint64_t n, uint32_t n_points) {
    // enough (sample block, point chunk) CTAs to fill the GPU even for a handful of samples
    const uint64_t blocks = (n + kThreads - 1) / kThreads;
    uint32_t want = (uint32_t)((2ull * (uint64_t)sm_count() + blocks - 1) / (blocks ? blocks : 1));
    uint32_t maxc = (n_points + kVolTile - 1) / kVolTile;
    if (want < 1) want = 1;
    return want < maxc ? want : (maxc ? maxc : 1);
}

ALPK_EXPORT int alpk_vol_int(const double* energy, const double* weight, uint64_t n, const alpk_prop_t* h_prop,
                             const double* nxm, const double* bden, const double* jac, uint32_t n_points,
                             double volume_over_n, double inv_mbm, float* decay_w32, float* scatter_w32,
                             double* decay_w64, double* scatter_w64, double* partials, void* stream) {
    if (!h_prop || !decay_w32 || !scatter_w32 || !partials || !nxm || !bden || !jac) { set_error("alpk_vol_int: null argument"); return 2; }
    if (n == 0) return 0;
    if (n_points == 0) { set_error("alpk_vol_int: no geometry points"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t n_chunks = alpk_vol_int_chunks(n, n_points);
    uint32_t chunk = (n_points + n_chunks - 1) / n_chunks;
    chunk = ((chunk + kVolTile - 1) / kVolTile) * kVolTile;
    const uint32_t used = (n_points + chunk - 1) / chunk;
    dim3 grid((unsigned)((n + kThreads - 1) / kThreads), used);
    PropConst c;
    c.ma = h_prop->ma; c.ma2 = h_prop->ma2; c.width = h_prop->width; c.rescale = h_prop->rescale;
    c.neg_dist_mbm = h_prop->neg_dist_mbm; c.neg_len_mbm = h_prop->neg_len_mbm; c.geom = 1.0; c.geom32 = 1.0f;
    c.apply_geom = 0; c.geom_f64 = 0; c.use_tof = 0; c.tof_light = 0; c.det_dist = h_prop->det_dist; c.timing_window = 0;
    c.fast = h_prop->fast; c.k1 = h_prop->k1; c.k2 = h_prop->k2;
    prop_derive(c);
    if (h_prop->fast) vol_int_kernel<true><<<grid, kThreads, 0, st>>>(energy, n, nxm, bden, jac, n_points, chunk, c.ma, c.ma2, c.width, inv_mbm, partials);
    else vol_int_kernel<false><<<grid, kThreads, 0, st>>>(energy, n, nxm, bden, jac, n_points, chunk, c.ma, c.ma2, c.width, inv_mbm, partials);
    if (check_launch("alpk_vol_int")) return 1;
    const int g2 = grid_for(n, kThreads, 8);
    if (h_prop->fast) vol_int_finish_kernel<true><<<g2, kThreads, 0, st>>>(energy, weight, n, partials, used, volume_over_n, c, decay_w32, scatter_w32, decay_w64, scatter_w64);
    else vol_int_finish_kernel<false><<<g2, kThreads, 0, st>>>(energy, weight, n, partials, used, volume_over_n, c, decay_w32, scatter_w32, decay_w64, scatter_w64);
    return check_launch("alpk_vol_int(finish)");
}
