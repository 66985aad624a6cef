// Row-aligned variant of the per-sample chain kernel for the production configuration: fast arithmetic, free-running
// Philox, even n_samples, rows of at least kRowsMinSamples samples.
//
//   Work unit = (row, chunk of the row's sample pairs), owned by ONE WARP: a unit never straddles two rows, so
//     * the row record is loaded once per unit into registers (warp-uniform __ldg, no shared-memory staging) -- the
//       tile loop of chain_kernel needs two __syncthreads per tile and a per-pair row select;
//     * the Philox counter of a pair is its pair index within the row and the row id is a per-unit constant;
//     * there is no CTA-level synchronisation in the loop at all: warps drift apart, which spreads their integer
//       (Philox) and FP64 phases over the sub-partition's issue slots.
//   Units are handed out dynamically (one atomicAdd per unit, fetched one unit ahead so its latency is hidden), which
//   removes the tail of a static assignment (measured: 0.508 ms static vs 0.472 ms dynamic on the C2 chain).
//   The event rate stays deterministic all the same: every unit's sum is formed in a fixed order (lane-local adds in
//   iteration order, then a fixed shuffle tree) and stored per UNIT; the finalize kernel adds the unit sums in unit order.
//   Which warp happened to run a unit does not enter the result, nor does the grid size.
//
// The per-sample arithmetic is the same inline code as chain_kernel's fast path (Prod::sample_fast, propagate_try_fast,
// decay_event_weight_coupled), and sample s of a row uses the same Philox counter / words, so both kernels produce the
// same bits for every per-sample output; only the order of the rate sum differs.
//
// Fused weighted histogram of E_a (HIST): np.histogram(E_a, edges, weights = weight of the last fused stage).  Each warp
// owns `hist_cols` private COLUMNS of the bins in shared memory (cell = bins[warp][bin][col]); K = 32 / hist_cols lanes share
// a column and take turns in K warp-synchronous phases, each lane doing plain read-modify-writes on its column: no atomics,
// no shuffles, no branches (lanes whose turn it is not, and samples outside the edges, add into a per-lane trash cell).
// The bin comes from a float spacing guess confirmed against the bin's edge pair (one 16-byte shared load, common.cuh
// find_bin_pairs).  hist_cols is the largest power of two whose bins fit the CTA's shared-memory budget (50 bins: 16 columns,
// K = 2).  Histograms with too many bins for four columns per warp keep ONE copy of the bins per CTA and use shared-memory
// atomics (collisions are rare there).  The CTA's bins are merged into the output with one global atomic per non-empty bin,
// so a spectrum is reproducible to rounding only (the rate sum stays bit-reproducible).  The match.any-aggregated adds of
// chain_kernel measured 1.24 ms in this kernel (50 bins), the first, branchy version of the columns 0.67 ms.
#pragma once
#include <math_constants.h>
#include <stdlib.h>

#include "chain_kernel.cuh"

namespace alpk {

#ifndef ALPK_ROWS_MIN_BLOCKS
#define ALPK_ROWS_MIN_BLOCKS 2             // 2 x 256 threads / SM: 128 registers per thread (3 blocks / 80 registers: 2 % slower)
#endif
#ifndef ALPK_ROWS_THREADS
#define ALPK_ROWS_THREADS 256
#endif
#ifndef ALPK_ROWS_UNROLL
#define ALPK_ROWS_UNROLL 1
#endif
#ifndef ALPK_ROWS_CHUNK_PAIRS
#define ALPK_ROWS_CHUNK_PAIRS 512          // target pairs per unit (16 warp iterations); 256 / 1024 measured 1-2 % slower
#endif
#ifndef ALPK_ROWS_STCS
#define ALPK_ROWS_STCS 1
#endif
#if ALPK_ROWS_STCS
#define ALPK_ROWS_ST(ptr, val) __stcs(ptr, val)
#else
#define ALPK_ROWS_ST(ptr, val) (*(ptr) = (val))
#endif
constexpr int kRowsThreads = ALPK_ROWS_THREADS;
constexpr int kRowsUnroll = ALPK_ROWS_UNROLL;
constexpr int kRowsHistSmem = 88 * 1024;    // dynamic shared memory of a HIST launch: 2 CTAs x (16 KB exp table + this) per SM
constexpr uint32_t kRowsMinSamples = 256;   // shorter rows stay with the tile kernel (units would be mostly lane padding)
// layout of the caller's scratch (>= ALPK_SCRATCH_DOUBLES doubles): [0, 2 kMaxPartials) block partials of the other kernels,
// then the dynamic-unit counter, then the per-unit rate sums of this kernel
constexpr uint64_t kRowsCounterOffset = 2 * kMaxPartials;      // [+0] unit counter (u64), [+1] arrival counter of the finalize kernel
constexpr uint64_t kRowsUnitSumOffset = 2 * kMaxPartials + 16;
constexpr uint64_t kRowsMaxUnits = ALPK_SCRATCH_DOUBLES - kRowsUnitSumOffset;

struct RowsGeom {
    uint32_t chunks_per_row;     // units per row
    uint32_t pairs_per_chunk;    // pairs per unit (the last unit of a row may hold fewer)
    uint32_t pairs_per_row;      // n_samples / 2
    uint32_t n_units;            // n_rows * chunks_per_row (<= kRowsMaxUnits)
    unsigned long long* counter; // dynamic unit counter (zeroed by the launcher), or NULL: static assignment
    double* unit_sums;           // [n_units] rate sum of every unit (STAGE 2)
    uint32_t hist_cols;          // HIST: bin columns per warp (32 .. 2), or 0: one copy of the bins per CTA, shared atomics
};

template <class Prod, int STAGE, int OUT, bool SHORT, int ROUNDS, bool HIST = false>
__global__ void __launch_bounds__(kRowsThreads, ALPK_ROWS_MIN_BLOCKS)
chain_rows_kernel(const ChainArgs<Prod> a, const RowsGeom geo) {
    using namespace alp;
    constexpr int NP = Prod::NP;
    __shared__ double s_tab[alp::kExpTabSize];
    extern __shared__ double2 s_rows_dyn[];          // HIST: edge pairs [nb], edges [n_edges], then the bins (hist_rows_doubles)
    for (int i = threadIdx.x; i < alp::kExpTabSize; i += blockDim.x) s_tab[i] = kExpTabDev[i];
    const uint32_t lane = threadIdx.x & 31u;
    const int n_hbins = HIST ? a.n_edges - 1 : 0;
    const double2* const s_pairs = s_rows_dyn;
    double* const s_edges = reinterpret_cast<double*>(s_rows_dyn) + 2 * n_hbins;
    double* const s_bins = s_edges + (HIST ? ((a.n_edges + 1) & ~1) : 0);
    const uint32_t hcols = HIST ? geo.hist_cols : 0u;             // columns per warp
    const uint32_t hphases = hcols ? 32u / hcols : 0u;            // lanes per column = warp-synchronous phases
    const uint32_t hwarp = HIST ? (uint32_t)n_hbins * hcols + 32u : 0u;   // cells per warp: bins x columns + one trash cell per lane
    const int n_hcells = HIST ? (hcols ? (int)hwarp * (kRowsThreads / 32) : n_hbins) : 0;
    double* my_col = s_bins;
    double* my_trash = s_bins;
    BinGuessF guess{0, 0.f, 0.f, 0.0, 0.0};
    if (HIST) {
        for (int i = threadIdx.x; i < a.n_edges; i += blockDim.x) s_edges[i] = __ldg(a.hist_edges + i);
        for (int i = threadIdx.x; i < n_hbins; i += blockDim.x) {
            const double hi = __ldg(a.hist_edges + i + 1);
            s_rows_dyn[i] = make_double2(__ldg(a.hist_edges + i), i == n_hbins - 1 ? nextafter(hi, CUDART_INF) : hi);
        }
        for (int i = threadIdx.x; i < n_hcells; i += blockDim.x) s_bins[i] = 0.0;
        __syncthreads();
        guess = to_float_guess(detect_spacing(s_edges, a.n_edges), s_edges, a.n_edges);
        if (hcols) {
            my_col = s_bins + (size_t)(threadIdx.x >> 5) * hwarp + lane / hphases;
            my_trash = s_bins + (size_t)(threadIdx.x >> 5) * hwarp + (uint32_t)n_hbins * hcols + lane;
        }
    }
    __syncthreads();
    const uint32_t n_units = geo.n_units;
    const uint32_t n_warps = gridDim.x * (kRowsThreads / 32);

    uint32_t unit = blockIdx.x * (kRowsThreads / 32) + (threadIdx.x >> 5);
    // the first n_warps units are assigned statically; with a counter the further ones are handed out dynamically,
    // fetched one unit ahead so that the atomic's latency is hidden behind the current unit
    const bool dynamic = geo.counter != nullptr;
    uint64_t next = (uint64_t)unit + n_warps;
    if (dynamic) {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(geo.counter, 1ull);
        next = n_warps + __shfl_sync(0xffffffffu, t, 0);
    }
    while (unit < n_units) {
        unsigned long long fetched = 0;
        if (dynamic && lane == 0 && next < n_units) fetched = atomicAdd(geo.counter, 1ull);
        const uint32_t r = unit / geo.chunks_per_row;
        const uint32_t c = unit - r * geo.chunks_per_row;
        const uint32_t p_begin = c * geo.pairs_per_chunk;
        const uint32_t p_stop = min(p_begin + geo.pairs_per_chunk, geo.pairs_per_row);
        double row[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) row[k] = __ldg(a.table + (uint64_t)r * NP + k);
        const uint32_t rid = a.row_ids ? __ldg(a.row_ids + r) : r;
        // per-lane output cursors of the unit, advanced by one warp trip (32 pairs) per iteration: two adds per array instead
        // of rebuilding every address from the sample index
        const uint64_t i_first = (uint64_t)r * a.n_samples + 2ull * (p_begin + lane);
        constexpr bool kAll = OUT == kOutStandard || OUT == kOutNoEvents;   // pointers known non-null (d64 / s64 known null)
        constexpr bool kEvw = OUT == kOutStandard;                            // event_w known non-null
        double2* pe = (kAll || a.out_e) ? reinterpret_cast<double2*>(a.out_e + i_first) : nullptr;
        double2* pf = (kAll || a.out_f) ? reinterpret_cast<double2*>(a.out_f + i_first) : nullptr;
        float2* pd32 = (STAGE >= 1 && (kAll || a.d32)) ? reinterpret_cast<float2*>(a.d32 + i_first) : nullptr;
        float2* ps32 = (STAGE >= 1 && (kAll || a.s32)) ? reinterpret_cast<float2*>(a.s32 + i_first) : nullptr;
        double2* pd64 = (STAGE >= 1 && !kAll && a.d64) ? reinterpret_cast<double2*>(a.d64 + i_first) : nullptr;
        double2* ps64 = (STAGE >= 1 && !kAll && a.s64) ? reinterpret_cast<double2*>(a.s64 + i_first) : nullptr;
        double2* pev = (STAGE >= 2 && (kEvw || (OUT == kOutGeneric && a.event_w))) ? reinterpret_cast<double2*>(a.event_w + i_first) : nullptr;
        double acc = 0.0;

#pragma unroll kRowsUnroll
        // (HIST: warp-uniform trip count -- the column phases synchronise the warp --, lanes past the end are masked by `in`)
        for (uint32_t p = p_begin + lane; HIST ? (p - lane < p_stop) : (p < p_stop); p += 32u) {
            const bool in = !HIST || p < p_stop;
            // sample s: counter s>>1 = pair index, words (x,y) / (z,w)
            const Philox4 q = philox4x32_rk<ROUNDS>(p, 0u, rid, Prod::kTag, a.rk);
            double u[2], e[2], f[2], d64v[2], s64v[2], evv[2];
            float d32v[2], s32v[2];
            u[0] = u53(q.x, q.y);
            u[1] = u53(q.z, q.w);
#pragma unroll
            for (int j = 0; j < 2; ++j) e[j] = Prod::sample_fast(u[j], row, a.g, f[j], s_tab);
            if (STAGE >= 1) {
                bool ok = true;
#pragma unroll
                for (int j = 0; j < 2; ++j) ok &= propagate_try_fast<SHORT, true>(e[j], f[j], a.prop, d64v[j], s64v[j], s_tab);
                if (!ok) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) propagate_sample_fast(e[j], f[j], a.prop, d64v[j], s64v[j]);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    d32v[j] = prop_cast_fast(d64v[j], a.prop);
                    s32v[j] = prop_cast_fast(s64v[j], a.prop);
                }
            }
            if (STAGE >= 2) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    evv[j] = decay_event_weight_plain(e[j], d32v[j], a.ev);
                    if (HIST) acc += in ? evv[j] : 0.0; else acc += evv[j];
                }
            }
            if (HIST) {
                // np.histogram(E_a, bins=edges, weights=<weight of the last fused stage>)
                int hb[2];
                double hw[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    hw[j] = STAGE >= 2 ? evv[j] : (STAGE == 1 ? (double)d32v[j] : f[j]);
                    hb[j] = find_bin_pairs(e[j], s_pairs, n_hbins, guess);
                }
                if (__any_sync(0xffffffffu, (hb[0] == -2) | (hb[1] == -2))) {        // rare: the spacing guess missed
#pragma unroll
                    for (int j = 0; j < 2; ++j) if (hb[j] == -2) hb[j] = find_bin(e[j], s_edges, a.n_edges);
                }
                if (!in) hb[0] = hb[1] = -1;
                if (hcols) {
                    // both samples in one bin: a single add, so that the two read-modify-writes below never alias
                    const bool same = hb[1] == hb[0];
                    hw[0] += same ? hw[1] : 0.0;
                    double* const c0 = hb[0] >= 0 ? my_col + (uint32_t)hb[0] * hcols : my_trash;
                    double* const c1 = (hb[1] >= 0 && !same) ? my_col + (uint32_t)hb[1] * hcols : my_trash;
                    // read-modify-write of the lane's column cells; lanes whose turn it is not add into their trash cell
                    auto phase = [&](bool act) {
                        double* const q0 = act ? c0 : my_trash;
                        double* const q1 = act ? c1 : my_trash;
                        const double v0 = *q0, v1 = *q1;
                        *q0 = v0 + hw[0];
                        *q1 = v1 + hw[1];
                    };
                    if (hphases == 1) {                              // a private column per lane: no turns (warp-uniform branches)
                        phase(true);
                    } else if (hphases == 2) {
                        phase((lane & 1u) == 0u);
                        __syncwarp();
                        phase((lane & 1u) == 1u);
                        __syncwarp();
                    } else {
                        for (uint32_t ph = 0; ph < hphases; ++ph) {
                            phase((lane & (hphases - 1u)) == ph);
                            __syncwarp();
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 2; ++j) if (hb[j] >= 0) atomicAdd(s_bins + hb[j], hw[j]);
                }
            }
            if (OUT != kOutNone && in) {
                // write-once outputs: streaming stores (evict-first in L2)
                if (kAll || pe) { ALPK_ROWS_ST(pe, make_double2(e[0], e[1])); pe += 32; }
                if (kAll || pf) { ALPK_ROWS_ST(pf, make_double2(f[0], f[1])); pf += 32; }
                if (STAGE >= 1) {
                    if (kAll || pd32) { ALPK_ROWS_ST(pd32, make_float2(d32v[0], d32v[1])); pd32 += 32; }
                    if (kAll || ps32) { ALPK_ROWS_ST(ps32, make_float2(s32v[0], s32v[1])); ps32 += 32; }
                    if (!kAll && pd64) { ALPK_ROWS_ST(pd64, make_double2(d64v[0], d64v[1])); pd64 += 32; }
                    if (!kAll && ps64) { ALPK_ROWS_ST(ps64, make_double2(s64v[0], s64v[1])); ps64 += 32; }
                }
                if (STAGE >= 2 && (kEvw || (OUT == kOutGeneric && pev))) { ALPK_ROWS_ST(pev, make_double2(evv[0], evv[1])); pev += 32; }
            }
        }
        if (STAGE >= 2) {                            // the unit's rate sum: fixed order whichever warp ran the unit
            acc = warp_sum(acc);
            if (lane == 0) geo.unit_sums[unit] = acc;
        }
        const uint64_t cur_next = next;
        // (a warp that did not fetch -- next was already past the end -- ends with unit >= n_units)
        next = dynamic ? (cur_next < n_units ? n_warps + __shfl_sync(0xffffffffu, fetched, 0) : cur_next) : cur_next + n_warps;
        unit = cur_next < n_units ? (uint32_t)cur_next : n_units;
    }
    if (HIST) {
        // the CTA's bins: warp copies and columns added in a fixed order, one global atomic per non-empty bin
        __syncthreads();
        const uint32_t per_bin = hcols ? hcols : 1u, copies = hcols ? kRowsThreads / 32 : 1;
        for (int i = threadIdx.x; i < n_hbins; i += blockDim.x) {
            double v = 0.0;
            for (uint32_t w = 0; w < copies; ++w)
                for (uint32_t c = 0; c < per_bin; ++c) v += s_bins[(size_t)w * hwarp + (size_t)i * per_bin + c];
            if (v != 0.0) atomicAdd(a.hist + i, v);
        }
    }
}

// Fixed-order sum of the unit sums: CTA b adds its contiguous block of units (thread-strided, then the block tree), the last
// CTA to finish adds the block sums in block order.  One launch; the result depends only on (n, grid, block size).
static __global__ void __launch_bounds__(256)
finalize_units_kernel(const double* __restrict__ sums, uint32_t n, double* __restrict__ partials, unsigned int* done,
                      double* __restrict__ out) {
    __shared__ double red[32];
    __shared__ bool s_last;
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * per, hi = min(lo + per, n);
    double v = 0.0;
    for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) v += sums[i];
    v = block_sum(v, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = v;
        __threadfence();
        s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double t = 0.0;
        for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += reinterpret_cast<const volatile double*>(partials)[i];
        t = block_sum(t, red);
        if (threadIdx.x == 0) { *out = t; *done = 0u; }
    }
}

// Geometry of the units for a given row length: equal-sized units of whole warp iterations, at most kRowsMaxUnits of them.
inline RowsGeom rows_geometry(uint64_t n_rows, uint32_t n_samples, double* scratch, bool dynamic) {
    RowsGeom g;
    g.pairs_per_row = n_samples / 2u;
    uint64_t target = ALPK_ROWS_CHUNK_PAIRS;
    for (;;) {
        g.chunks_per_row = (uint32_t)((g.pairs_per_row + target - 1) / target);
        uint32_t ppc = (g.pairs_per_row + g.chunks_per_row - 1) / g.chunks_per_row;
        ppc = (ppc + 31u) & ~31u;
        g.pairs_per_chunk = ppc;
        g.chunks_per_row = (g.pairs_per_row + ppc - 1) / ppc;
        if (n_rows * (uint64_t)g.chunks_per_row <= kRowsMaxUnits || g.chunks_per_row == 1) break;
        target *= 2;                                 // very many rows: larger units so the unit sums fit the scratch
    }
    g.n_units = (uint32_t)(n_rows * (uint64_t)g.chunks_per_row);
    g.counter = (dynamic && scratch) ? reinterpret_cast<unsigned long long*>(scratch + kRowsCounterOffset) : nullptr;
    g.unit_sums = scratch ? scratch + kRowsUnitSumOffset : nullptr;
    return g;
}

inline int chain_rows_mode() {
    // ALPK_CHAIN_ROWS=0 keeps every launch on the tile kernel, 2 = row kernel with static unit assignment (A/B timing and
    // the kernel-vs-kernel bit-identity test); default 1.  Read per launch so a test can flip it.
    const char* e = getenv("ALPK_CHAIN_ROWS");
    return e ? atoi(e) : 1;
}

// Dynamic shared memory of a HIST launch, in doubles: edge pairs [nb][2], edges [n_edges] (padded to even), then per warp
// nb x cols bin cells + 32 trash cells -- or, cols == 0, one copy of the nb bins for the CTA.
inline size_t hist_rows_doubles(int n_edges, uint32_t cols) {
    const size_t nb = (size_t)n_edges - 1;
    return 2 * nb + (((size_t)n_edges + 1) & ~(size_t)1) + (cols ? (nb * cols + 32) * (kRowsThreads / 32) : nb);
}

inline int hist_rows_mode() {
    // ALPK_HIST_ROWS=0 keeps the fused-histogram launches on the tile kernel (A/B timing, kernel-vs-kernel test); default 1
    const char* e = getenv("ALPK_HIST_ROWS");
    return e ? atoi(e) : 1;
}

template <class Prod>
int launch_chain_rows(const char* what, ChainArgs<Prod>& a, int stage, bool standard, bool noevents, bool none,
                      double* rate, double* scratch, cudaStream_t st) {
    const int mode = chain_rows_mode();
    if (!mode || a.n_samples < kRowsMinSamples) return -1;
    // the sample loop is compiled without the rare-mode tests: FP32 transcendentals and an np.float64 exposure stay with the
    // tile kernel
    if ((stage >= 1 && a.prop.fast != 1) || (stage >= 2 && a.ev.days_f64)) return -1;
    RowsGeom geo = rows_geometry(a.n_rows, a.n_samples, scratch, mode != 2);
    if ((uint64_t)a.n_rows * geo.chunks_per_row > kRowsMaxUnits) return -1;          // (more rows than unit slots: tile kernel)
    size_t smem = 0;
    geo.hist_cols = 0;
    if (a.hist) {
        // the fused spectrum of this kernel is the event-weighted one (stage 2: what flux.chain() asks for); histograms of the
        // production flux or of the decay weights alone stay with the tile kernel
        if (!hist_rows_mode() || stage != 2) return -1;
        uint32_t cols = 32;
        while (cols >= 4 && hist_rows_doubles(a.n_edges, cols) * sizeof(double) > (size_t)kRowsHistSmem) cols >>= 1;
        geo.hist_cols = cols >= 4 ? cols : 0;
        smem = hist_rows_doubles(a.n_edges, geo.hist_cols) * sizeof(double);
        if (smem > (size_t)kRowsHistSmem) return -1;                                  // (too many bins: tile kernel / alpk_hist)
    }
    // (the unit counter and the finalize kernel's arrival counter sit next to each other: one 16-byte memset)
    if (scratch) cudaMemsetAsync(scratch + kRowsCounterOffset, 0, 16, st);
    uint64_t blocks = ((uint64_t)geo.n_units + (kRowsThreads / 32) - 1) / (kRowsThreads / 32);
    const uint64_t cap = (uint64_t)sm_count() * ALPK_ROWS_MIN_BLOCKS;
    if (blocks > cap) blocks = cap;
    const int grid = (int)blocks;
    const bool sh = stage >= 1 && a.prop.short_lived;
#define ALPK_ROWS_LAUNCH_K(S, O, SH, R, H) {                                                                              \
        auto kfn = chain_rows_kernel<Prod, S, O, SH, R, H>;                                                                \
        if (H) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowsHistSmem);                      \
        kfn<<<grid, kRowsThreads, smem, st>>>(a, geo); }
#define ALPK_ROWS_LAUNCH_R(S, O, R, H) { if (sh) ALPK_ROWS_LAUNCH_K(S, O, true, R, H) else ALPK_ROWS_LAUNCH_K(S, O, false, R, H) }
#define ALPK_ROWS_LAUNCH(S, O, H) { if (a.rounds == 7) ALPK_ROWS_LAUNCH_R(S, O, 7, H) else ALPK_ROWS_LAUNCH_R(S, O, 10, H) }
#define ALPK_ROWS_CASE(S)                                                         \
    if (standard) ALPK_ROWS_LAUNCH(S, kOutStandard, false)                        \
    else if (S == 2 && noevents) ALPK_ROWS_LAUNCH(S, kOutNoEvents, false)         \
    else if (none) ALPK_ROWS_LAUNCH(S, kOutNone, false)                           \
    else ALPK_ROWS_LAUNCH(S, kOutGeneric, false)
    if (a.hist) {                                  // (stage 2 only: checked above)
        if (none) ALPK_ROWS_LAUNCH(2, kOutNone, true) else ALPK_ROWS_LAUNCH(2, kOutGeneric, true)
    } else if (stage == 0) { ALPK_ROWS_CASE(0) } else if (stage == 1) { ALPK_ROWS_CASE(1) } else { ALPK_ROWS_CASE(2) }
#undef ALPK_ROWS_LAUNCH
#undef ALPK_ROWS_LAUNCH_R
#undef ALPK_ROWS_LAUNCH_K
#undef ALPK_ROWS_CASE
    if (check_launch(what)) return 1;
    if (stage == 2) {
        if (geo.n_units <= (uint32_t)kMaxPartials) return finalize_sum(geo.unit_sums, (int)geo.n_units, rate, st);
        const int fgrid = (int)min((uint64_t)sm_count(), ((uint64_t)geo.n_units + 511) / 512);
        finalize_units_kernel<<<fgrid, 256, 0, st>>>(geo.unit_sums, geo.n_units, scratch,
                                                     reinterpret_cast<unsigned int*>(scratch + kRowsCounterOffset + 1), rate);
        return check_launch("finalize_units");
    }
    return 0;
}

}  // namespace alpk
