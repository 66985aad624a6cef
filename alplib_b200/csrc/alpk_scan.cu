// libalpk: batched (m_a, g) sensitivity scan -- the whole coupling axis and mass axis in one launch grid.
//
// Reference pattern (README.md:199-226, examples/ex4_na64.py:30-40, fit.py:12-58): for every m_a one simulate(), for
// every (m_a, g) one propagate(new_coupling=g) + decays().  propagate(new_coupling) only rescales the production
// weights by (g/g0)^2 and recomputes the width (fluxes.py:153-158, 235-240), so the kernel
//   blockIdx.y = m_a index, blockIdx.x = tile of that m_a's samples, threads = coupling index
//   stage 1: the CTA regenerates its tile of samples (same Philox keys as the per-point API) and parks the
//            g-independent part of propagate (1/p_a or v_a and boost, weight, threshold flag) in shared memory;
//   stage 2: every thread owns one coupling and walks the tile (shared-memory broadcast reads), accumulating
//            days*w32*heaviside exactly as fluxes.py:61-65,159-162 + generators.py:88-92 evaluate it;
//   partial sums [m_a][tile][g] are reduced by a second kernel in fixed order (deterministic).
#include <string.h>

#include "../../include/alpk.h"
#include "common.cuh"

using namespace alp;

namespace alpk {

constexpr int kScanTile = 1024;      // samples per CTA tile
constexpr int kScanThreads = 256;

struct ScanGeom {
    double neg_dist_mbm, neg_len_mbm, geom;
    float geom32;
    int apply_geom, geom_f64, fast;
    EventConst ev;
};

// ---- per-(m_a, row) tables ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
scan_compton_rows_kernel(const double* __restrict__ eg, const double* __restrict__ phi, uint32_t n_rows,
                         const double* __restrict__ log10_e, const double* __restrict__ log10_xs, int n_table,
                         const double* __restrict__ ma_grid, const double* __restrict__ ma2_grid, int n_ma,
                         double aa, double z, double n_samples, double* __restrict__ tables) {
    __shared__ double s_e[512], s_xs[512];
    for (int i = threadIdx.x; i < n_table; i += blockDim.x) { s_e[i] = __ldg(log10_e + i); s_xs[i] = __ldg(log10_xs + i); }
    __syncthreads();
    const uint64_t total = (uint64_t)n_ma * n_rows;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t im = (uint32_t)(i / n_rows), r = (uint32_t)(i - (uint64_t)im * n_rows);
        const double e = __ldg(eg + r);
        ComptonGlobals g{__ldg(ma_grid + im), __ldg(ma2_grid + im), aa, z, n_samples};
        double row[CB_NP];
        const double s = (2 * kME) * e + kME * kME;                         // fluxes.py:214
        const double thr = kME + g.ma;
        if (s < thr * thr) {                                                // :215 row produces nothing for this m_a
#pragma unroll
            for (int k = 0; k < CB_NP; ++k) row[k] = 0.0;
            row[CB_EG] = -1.0;
        } else {
            compton_bin_params(e, __ldg(phi + r), abs_sigma_mev(e, s_e, s_xs, n_table), g, row);
        }
#pragma unroll
        for (int k = 0; k < CB_NP; ++k) tables[i * CB_NP + k] = row[k];
    }
}

// Primakoff: one sample per row; table row = {E_a (or -1), production weight}
__global__ void __launch_bounds__(kThreads)
scan_primakoff_rows_kernel(const double* __restrict__ eg, const double* __restrict__ phi, uint32_t n_rows,
                           const double* __restrict__ log10_e, const double* __restrict__ log10_xs, int n_table,
                           const double* __restrict__ ma_grid, int n_ma, double pref, double r02,
                           double* __restrict__ tables) {
    __shared__ double s_e[512], s_xs[512];
    for (int i = threadIdx.x; i < n_table; i += blockDim.x) { s_e[i] = __ldg(log10_e + i); s_xs[i] = __ldg(log10_xs + i); }
    __syncthreads();
    const uint64_t total = (uint64_t)n_ma * n_rows;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t im = (uint32_t)(i / n_rows), r = (uint32_t)(i - (uint64_t)im * n_rows);
        const double e = __ldg(eg + r), ma = __ldg(ma_grid + im);
        if (e < ma) {                                                       // fluxes.py:136
            tables[i * 2] = -1.0; tables[i * 2 + 1] = 0.0;
        } else {
            const double br = primakoff_sigma(e, ma, pref, r02) / abs_sigma_mev(e, s_e, s_xs, n_table);
            tables[i * 2] = e; tables[i * 2 + 1] = __ldg(phi + r) * br;
        }
    }
}

// ---- the scan kernel ---------------------------------------------------------------------------------------------------
struct ScanArgs {
    const double* tables;        // [n_ma][n_rows][NP]
    uint32_t n_rows, n_samples; int n_ma, n_g;
    const double* ma_grid; const double* ma2_grid;
    const double* width;         // [n_ma][n_g]
    const double* rescale;       // [n_g]
    double aa, z; uint64_t seed; uint32_t row_offset; int rounds;
    ScanGeom geo;
    double* partials;            // [n_ma][n_tiles][n_g]  (n_tiles = tiles of this launch: the tile axis is walked in chunks)
    uint32_t n_tiles;
    uint32_t tile0;              // first tile of this launch
};

template <bool COMPTON, bool FAST>
__global__ void __launch_bounds__(kScanThreads)
scan_kernel(const ScanArgs a) {
    __shared__ double s_t[kScanTile];       // fast: 1/p_a      faithful: v_a
    __shared__ double s_b[kScanTile];       //                  faithful: boost = E_a/m_a
    __shared__ double s_w[kScanTile];       // production weight (0 for dead samples)
    __shared__ double s_h[kScanTile];       // heaviside(E_a - threshold, 1)
    __shared__ double s_wh[kScanTile];      // fast path: weight with the (exact 0/1/NaN) threshold flag folded in
    constexpr int NP = COMPTON ? (int)CB_NP : 2;
    const int im = blockIdx.y;
    const uint32_t tile = a.tile0 + blockIdx.x;
    const double ma = __ldg(a.ma_grid + im), ma2 = __ldg(a.ma2_grid + im);
    const uint64_t n = (uint64_t)a.n_rows * a.n_samples;
    const uint64_t base = (uint64_t)tile * kScanTile;
    const int count = (int)((base + kScanTile <= n ? (uint64_t)kScanTile : n - base));
    const double* table = a.tables + (uint64_t)im * a.n_rows * NP;
    const ComptonGlobals cg{ma, ma2, a.aa, a.z, (double)a.n_samples};

    // ---- stage 1: regenerate the tile's samples; keep only those that can contribute (weight != 0 above threshold),
    //      compacted in their original order (block-wide exclusive scan), so stage 2 walks live samples only and the
    //      sums are bit-identical to walking the whole tile (the skipped terms are exact zeros) ------------------------
    __shared__ int s_warp[kScanThreads / 32];
    __shared__ int s_live;
    // exp_fast table: read through L1 from its global copy (16 KB; the five sample arrays already take 40 KB of the 48 KB
    // of static shared memory a CTA may declare)
    const double* const s_tab = nullptr;
    __shared__ unsigned long long s_tmax_bits;   // max over live samples of 1/p_a (non-finite / NaN -> +inf), as ordered bits
    __shared__ unsigned long long s_tmin_bits;   // min over live samples of 1/p_a; 0 if any live sample has a NaN 1/p_a or a non-finite weight
    __shared__ unsigned long long s_wmax_bits;   // max |weight| over live samples
    if (threadIdx.x == 0) { s_tmax_bits = 0ull; s_tmin_bits = 0x7ff0000000000000ull; s_wmax_bits = 0ull; }
    __syncthreads();
    unsigned long long my_tmax = 0ull, my_tmin = 0x7ff0000000000000ull, my_wmax = 0ull;
    constexpr int kPer = kScanTile / kScanThreads;                  // consecutive samples per thread
    double rt_[kPer], rb_[kPer], rw_[kPer], rh_[kPer];
    int nkeep = 0;
    unsigned keepmask = 0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        const int k = threadIdx.x * kPer + q;
        if (k >= count) continue;
        const uint64_t i = base + k;
        const uint32_t r = (uint32_t)(i / a.n_samples), s = (uint32_t)(i - (uint64_t)r * a.n_samples);
        const double* b = table + (size_t)r * NP;
        double ea = 0.0, w = 0.0;
        const bool live = __ldg(b) >= 0.0;
        if (live) {
            if (COMPTON) {
                double row[CB_NP];
#pragma unroll
                for (int c = 0; c < CB_NP; ++c) row[c] = __ldg(b + c);
                const double u = uniform_1(a.seed, kTagComptonEa, a.row_offset + r, s, a.rounds);
                ea = FAST ? compton_sample_fast(u, row, cg, w) : compton_sample(u, row, cg, w);
            } else {
                ea = __ldg(b); w = __ldg(b + 1);
            }
        }
        const double h = live ? heaviside(ea - a.geo.ev.threshold, 1.0) : 0.0;
        if (!((w * h != 0.0) || (w != w))) continue;                // exact zero contribution for every coupling
        const double d = ea * ea - ma2;
        double t, bo = 0.0;
        if (FAST) {
            // E_a <= m_a: the reference gives surv = exp(-inf) = 0 (v = 0) or NaN (E_a < m_a); 1/p_a = inf / NaN
            t = d > 1e-300 ? rsqrt_fast1(d) : (d > 0.0 ? rsqrt(d) : ((d == 0.0) ? INFINITY : NAN));   // as propagate_sample_fast
        } else {
            const double pa = sqrt(d);
            t = pa / ea;                 // v_a
            bo = ea / ma;                // boost
        }
        rt_[q] = t; rb_[q] = bo; rw_[q] = w; rh_[q] = h;      // (static register indexing; compacted on the store)
        {   // t >= 0 or NaN: non-negative doubles order like their bit patterns
            const unsigned long long tb = (t >= 0.0 && t < INFINITY) ? (unsigned long long)__double_as_longlong(t)
                                                                     : 0x7ff0000000000000ull;
            my_tmax = tb > my_tmax ? tb : my_tmax;
            const bool clean = (t >= 0.0) && (fabs(w) < INFINITY);         // (false for NaN t / NaN or infinite w)
            const unsigned long long lb = clean ? (unsigned long long)__double_as_longlong(t) : 0ull;
            my_tmin = lb < my_tmin ? lb : my_tmin;
            const unsigned long long wb = clean ? (unsigned long long)__double_as_longlong(fabs(w)) : 0ull;
            my_wmax = wb > my_wmax ? wb : my_wmax;
        }
        keepmask |= 1u << q;
        ++nkeep;
    }
    // exclusive scan of nkeep over the CTA (warp shuffle scan + per-warp totals)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = nkeep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[wid] = incl;
    if (my_tmax) atomicMax(&s_tmax_bits, my_tmax);
    if (my_tmin != 0x7ff0000000000000ull) atomicMin(&s_tmin_bits, my_tmin);
    if (my_wmax) atomicMax(&s_wmax_bits, my_wmax);
    __syncthreads();
    int woff = 0;
    for (int q = 0; q < wid; ++q) woff += s_warp[q];
    if (threadIdx.x == blockDim.x - 1) s_live = woff + incl;
    int pos = woff + incl - nkeep;
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
        if (keepmask & (1u << q)) {
            // s_wh: the weight with the threshold flag folded in for the fast loop -- h is exactly 1 for a kept sample
            // unless the weight is inf/NaN (w*0 = NaN, NaN*h = NaN), which the product reproduces
            s_t[pos] = rt_[q]; s_b[pos] = rb_[q]; s_w[pos] = rw_[q]; s_h[pos] = rh_[q];
            s_wh[pos] = rh_[q] == 1.0 ? rw_[q] : rw_[q] * rh_[q];
            ++pos;
        }
    }
    __syncthreads();
    const int n_live = s_live;

    // ---- stage 2: one coupling per WARP at a time, the 32 lanes stride over the tile's live samples (conflict-free
    //      shared-memory reads, every lane busy for any n_g), lane partials combined by a shuffle tree (fixed order) ----
    const ScanGeom& G = a.geo;
    const double tmax = __longlong_as_double((long long)s_tmax_bits);      // largest finite 1/p_a of the tile (or +inf)
    const double tmin = __longlong_as_double((long long)s_tmin_bits);      // smallest 1/p_a (0 if a live sample is not clean)
    const double wmax = __longlong_as_double((long long)s_wmax_bits);
    for (int j = wid; j < a.n_g; j += kScanThreads / 32) {
        const double width = __ldg(a.width + (size_t)im * a.n_g + j);
        const double resc = __ldg(a.rescale + j);
        const double k1 = width > 0.0 ? (G.neg_dist_mbm * ma) * width : 0.0;
        const double k2 = width > 0.0 ? (G.neg_len_mbm * ma) * width : 0.0;
        // same validity rule as propagate_try_fast / prop_derive
        const double kmax = fmax(fabs(k1), fabs(k2));
        const double tf = fmin(kmax > 0.0 ? 708.0 / kmax : INFINITY, 1e140);
        double acc = 0.0;
        if (FAST && G.fast == 1 && tmin > 0.0 &&
            ((!(width > 0.0) && tmax < INFINITY) || (width > 0.0 && k1 * tmin + log(fabs(resc) * wmax) < -105.0))) {
            // stable ALP (width <= 0, e.g. m_a < 2 m_e for the electron coupling: half of the C5 mass axis): the decay
            // probability 1 - exp(0) is exactly 0 for every (clean) sample, so is every term of the sum.
            // short-lived corner of the grid: every decay weight of the tile is below half the smallest float32 denormal
            // (2^-150 = e^-103.97), i.e. the reference's float32 cast (fluxes.py:64) makes it exactly 0 --
            //   |resc w surv dec| <= |resc| w_max exp(k1 t_min),  dec in [0, 1],  surv <= exp(k1 t_min) as k1 < 0 <= t_min <= t
            // -- so every term of this (tile, coupling) sum is an exact zero and the walk is skipped.
        } else if (FAST && G.fast == 1 && !G.ev.days_f64 && tmax <= tf) {
            // every sample of the tile is inside the branch-free domain: straight-line body, two samples in flight
#pragma unroll 2
            for (int k = lane; k < n_live; k += 32) {
                const double t = s_t[k];
                const double surv = exp_surv(k1 * t, s_tab);
                const double dec = 1 - exp_dec(k2 * t, s_tab);
                const double s64 = (resc * s_wh[k]) * surv;
                const double d64 = s64 * dec;                                         // fluxes.py:64
                const float d32 = (float)d64 * G.geom32;                              // prop_cast_fast (:159-162)
                acc += (double)(G.ev.days_sec32 * d32);                               // generators.py:90 (h folded into s_wh)
            }
        } else {
            for (int k = lane; k < n_live; k += 32) {
                const double w = s_w[k];
                double surv, dec;
                if (FAST && G.fast == 2) {               // optional FP32 transcendentals (see propagate_sample_fast)
                    const float tf32 = (float)s_t[k];
                    surv = (double)expf((float)k1 * tf32);
                    dec = (double)(-expm1f((float)k2 * tf32));
                } else if (FAST) {
                    const double t = s_t[k];
                    if (t <= tf) { surv = exp_surv(k1 * t, s_tab); dec = 1 - exp_dec(k2 * t, s_tab); }
                    else { surv = exp(k1 * t); dec = 1 - exp(k2 * t); }
                } else {
                    const double va = s_t[k];
                    const double tau = width > 0.0 ? s_b[k] / width : INFINITY;      // fluxes.py:51
                    surv = exp((G.neg_dist_mbm / va) / tau);                          // :61
                    dec = 1 - exp((G.neg_len_mbm / va) / tau);                        // :62
                }
                const double s64 = (resc * w) * surv;
                const double d64 = s64 * dec;                                         // :64
                float d32 = (float)d64;
                if (FAST) d32 = d32 * G.geom32;                                       // prop_cast_fast
                else if (G.apply_geom) d32 = G.geom_f64 ? (float)((double)d32 * G.geom) : d32 * G.geom32;   // :159-162
                const double base_w = G.ev.days_f64 ? G.ev.days_sec * (double)d32 : (double)(G.ev.days_sec32 * d32);
                acc += base_w * s_h[k];                                               // generators.py:90
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) a.partials[((size_t)im * a.n_tiles + blockIdx.x) * a.n_g + j] = acc;
    }
}

// Fixed-order sum over the tiles of one chunk; later chunks continue the running sum of the earlier ones (the order of the
// additions is the tile order whatever the chunking, so the map does not depend on the size of the partials buffer).
__global__ void scan_finalize_kernel(const double* __restrict__ partials, int n_ma, uint32_t n_tiles, int n_g,
                                     int accumulate, double* __restrict__ out) {
    const uint64_t total = (uint64_t)n_ma * n_g;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t im = (uint32_t)(i / n_g), j = (uint32_t)(i - (uint64_t)im * n_g);
        double acc = accumulate ? out[i] : 0.0;
        for (uint32_t t = 0; t < n_tiles; ++t) acc += partials[((size_t)im * n_tiles + t) * n_g + j];
        out[i] = acc;
    }
}

}  // namespace alpk

using namespace alpk;

ALPK_EXPORT uint32_t alpk_scan_tiles(uint32_t n_rows, uint32_t n_samples) {
    const uint64_t n = (uint64_t)n_rows * n_samples;
    return (uint32_t)((n + kScanTile - 1) / kScanTile);
}

ALPK_EXPORT int alpk_scan(int process, const double* e_gamma, const double* phi_gamma, uint32_t n_rows, uint32_t row_offset,
                          const double* log10_e, const double* log10_xs, int n_table,
                          const double* ma_grid, const double* ma2_grid, int n_ma,
                          const double* width, const double* rescale, int n_g,
                          double c0, double c1, double z, uint32_t n_samples, uint64_t seed, int rng_rounds,
                          const alpk_prop_t* h_geom, const alpk_event_t* h_ev,
                          double* tables, double* partials, uint32_t partial_tiles, double* out_rates, void* stream) {
    if (process != 0 && process != 1) { set_error("alpk_scan: process must be 0 (Compton) or 1 (Primakoff)"); return 2; }
    if (!h_geom || !h_ev || !tables || !partials || !out_rates || !ma_grid || !ma2_grid || !width || !rescale) {
        set_error("alpk_scan: null argument"); return 2;
    }
    if (n_table < 1 || n_table > 512) { set_error("alpk_scan: table size %d not in [1,512]", n_table); return 2; }
    if (h_geom->use_tof) { set_error("alpk_scan: time-of-flight cut is not supported in the batched scan"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n_ma <= 0 || n_g <= 0) return 0;
    if (process == 1) n_samples = 1;
    if (n_rows == 0 || n_samples == 0) {
        cudaMemsetAsync(out_rates, 0, sizeof(double) * (size_t)n_ma * n_g, st);
        return 0;
    }
    const uint64_t total_rows = (uint64_t)n_ma * n_rows;
    if (process == 0) {
        scan_compton_rows_kernel<<<grid_for(total_rows, kThreads, 8), kThreads, 0, st>>>(
            e_gamma, phi_gamma, n_rows, log10_e, log10_xs, n_table, ma_grid, ma2_grid, n_ma, c0, z, (double)n_samples, tables);
    } else {
        scan_primakoff_rows_kernel<<<grid_for(total_rows, kThreads, 8), kThreads, 0, st>>>(
            e_gamma, phi_gamma, n_rows, log10_e, log10_xs, n_table, ma_grid, n_ma, c0, c1, tables);
    }
    if (check_launch("alpk_scan(rows)")) return 1;

    ScanArgs a;
    memset(&a, 0, sizeof(a));
    a.tables = tables; a.n_rows = n_rows; a.n_samples = n_samples; a.n_ma = n_ma; a.n_g = n_g;
    a.ma_grid = ma_grid; a.ma2_grid = ma2_grid; a.width = width; a.rescale = rescale;
    a.aa = c0; a.z = z; a.seed = seed; a.row_offset = row_offset; a.rounds = rng_rounds == 7 ? 7 : 10;
    a.geo.neg_dist_mbm = h_geom->neg_dist_mbm; a.geo.neg_len_mbm = h_geom->neg_len_mbm; a.geo.geom = h_geom->geom;
    a.geo.geom32 = h_geom->geom32; a.geo.apply_geom = h_geom->apply_geom; a.geo.geom_f64 = h_geom->geom_f64;
    a.geo.fast = h_geom->fast;
    a.geo.ev.days_sec = h_ev->days_sec; a.geo.ev.days_sec32 = h_ev->days_sec32; a.geo.ev.days_f64 = h_ev->days_f64;
    a.geo.ev.threshold = h_ev->threshold;
    a.partials = partials;
    const uint32_t all_tiles = alpk_scan_tiles(n_rows, n_samples);
    if (partial_tiles == 0) { set_error("alpk_scan: partials buffer holds no tile"); return 2; }
    if (n_ma > 65535) { set_error("alpk_scan: at most 65535 masses per call"); return 2; }
    const bool fast = h_geom->fast != 0;
    // the tile axis in chunks of what the caller's partials buffer holds (scratch stays bounded for any sample count)
    for (uint32_t t0 = 0; t0 < all_tiles; t0 += partial_tiles) {
        a.tile0 = t0;
        a.n_tiles = all_tiles - t0 < partial_tiles ? all_tiles - t0 : partial_tiles;
        dim3 grid(a.n_tiles, (unsigned)n_ma);
        if (process == 0) {
            if (fast) scan_kernel<true, true><<<grid, kScanThreads, 0, st>>>(a);
            else scan_kernel<true, false><<<grid, kScanThreads, 0, st>>>(a);
        } else {
            if (fast) scan_kernel<false, true><<<grid, kScanThreads, 0, st>>>(a);
            else scan_kernel<false, false><<<grid, kScanThreads, 0, st>>>(a);
        }
        if (check_launch("alpk_scan")) return 1;
        scan_finalize_kernel<<<grid_for((uint64_t)n_ma * n_g, 256, 4), 256, 0, st>>>(partials, n_ma, a.n_tiles, n_g, t0 != 0,
                                                                                     out_rates);
        if (check_launch("alpk_scan(finalize)")) return 1;
    }
    return 0;
}
