// Generic per-sample kernel of the flux pipeline: production (+ fused propagate (+ fused decays)).
//
//   Prod supplies: NP (doubles per row record), kTag (Philox stream tag), Globals, uidx() (layout of injected
//   uniforms) and sample(u, row, globals, flux) -> E_a.  One uniform per sample.
//
//   Work decomposition: a CTA walks tiles of kThreads*2*kUnroll consecutive samples (grid = multiple of the SM
//   count); the per-row records a tile touches are staged in shared memory (rows are long, so a tile touches 1-2 of
//   them; for tiny n_samples the table is read through L1/L2 instead).  Each thread handles PAIRS of adjacent samples:
//   one Philox4x32-10 call (2 x 53-bit uniforms) feeds both and stores are 16-byte (f64) / 8-byte (f32) vectors, fully
//   coalesced.  Rates are reduced warp-shuffle -> shared -> one partial per CTA -> deterministic finalize kernel.
#pragma once
#include "common.cuh"

namespace alpk {

#ifndef ALPK_UNROLL
#define ALPK_UNROLL 16
#endif
#ifndef ALPK_MIN_BLOCKS
#define ALPK_MIN_BLOCKS 4
#endif
#ifndef ALPK_GRID_PER_SM
#define ALPK_GRID_PER_SM 8
#endif
#ifndef ALPK_INNER_UNROLL
#define ALPK_INNER_UNROLL 1
#endif
constexpr int kUnroll = ALPK_UNROLL;
constexpr int kInnerUnroll = ALPK_INNER_UNROLL;
constexpr int kTile = kThreads * 2 * kUnroll;

template <class Prod>
struct ChainArgs {
    const double* table; uint64_t n_rows; uint32_t n_samples; const uint32_t* row_ids;
    typename Prod::Globals g; const double* uniforms; uint64_t seed;
    alp::PhiloxKeys rk;                                       // key schedule of `seed`
    int rounds;                                               // Philox rounds of the sample stream: 10 or 7
    double* out_e; double* out_f;
    alp::PropConst prop; float* d32; float* s32; double* d64; double* s64;
    alp::EventConst ev; double* event_w; double* partials;
    uint32_t tile_elems;
    const double* hist_edges; int n_edges; double* hist;     // optional fused weighted histogram of E_a
    int hist_copies;                                          // kThreads/32: one copy of the bins per warp; 1: one per CTA
};

// Which per-sample arrays a launch writes: checked once on the host so the specialised kernels carry no null tests.
enum ChainOut : int {
    kOutGeneric = 0,      // any subset (null pointer = not written)
    kOutStandard = 1,     // exactly the reference's attributes of the stage: E_a, flux [, f32 decay/scatter weights [, event weights]]
    kOutNone = 2,         // nothing (rate / histogram only)
    kOutNoEvents = 3      // STAGE 2: E_a, flux, f32 decay/scatter weights but no per-sample event weights (decays() rate only)
};

// One pair of adjacent samples (i0 = base + local, i0 even relative to the tile).
//   PAIRED: n_samples is even, so the two samples always lie in the same row and share one Philox call (counter s>>1):
//           no row-straddle / lone-sample handling and no word selects.
//   CHECK:  the tile may be partial (have2 can be false).
// Row / sample / Philox row id of the first row(s) of a tile, computed once per tile for the LONG fast loop (rows longer
// than a tile: a tile touches at most two rows, split at sample `split` of the tile).
struct TileRows { uint32_t split, rid0, rid1; };

//   LONG:   (with PAIRED, full tiles) rows are longer than a tile: row / sample index and row id of a pair come from the
//           tile-level TileRows with three instructions instead of the general index arithmetic and a row-id load.
template <class Prod, bool INJECT, int STAGE, bool FAST, bool HIST, bool PAIRED, int OUT, bool SHORT, bool CHECK, bool LONG = false>
__device__ __forceinline__ void chain_pair(const ChainArgs<Prod>& a, const double* rows, const double* s_tab,
                                           const double* s_edges, double* s_bins, uint64_t r0, uint32_t off0,
                                           uint64_t base, uint32_t local, uint32_t lim, bool long_rows, double& acc,
                                           const TileRows& tr = TileRows{0u, 0u, 0u},
                                           const BinGuess& guess = BinGuess{0, 0.0, 0.0}) {
    using namespace alp;
    constexpr int NP = Prod::NP;
    const uint64_t i0 = base + local;
    const bool have2 = (PAIRED && !CHECK) || (local + 1 < lim);
    double e[2], f[2], u[2];
    float d32v[2], s32v[2];
    double d64v[2], s64v[2], evv[2];
    uint32_t rr[2], ss[2];
    uint32_t rid_long = 0;
    if (LONG) {
        const bool second = local >= tr.split;
        rr[0] = rr[1] = second ? 1u : 0u;
        ss[0] = second ? local - tr.split : off0 + local;
        ss[1] = ss[0] + 1u;
        rid_long = second ? tr.rid1 : tr.rid0;
    } else {
        // (row, sample) of the pair: one integer division at most; none when rows are longer than a tile
        const uint32_t off = off0 + local;
        if (long_rows) { rr[0] = off >= a.n_samples ? 1u : 0u; ss[0] = off - (rr[0] ? a.n_samples : 0u); }
        else { rr[0] = off / a.n_samples; ss[0] = off - rr[0] * a.n_samples; }
        if (PAIRED) { rr[1] = rr[0]; ss[1] = ss[0] + 1u; }
        else {
            const bool wrap = ss[0] + 1u == a.n_samples;
            rr[1] = have2 ? (wrap ? rr[0] + 1u : rr[0]) : rr[0];
            ss[1] = have2 ? (wrap ? 0u : ss[0] + 1u) : ss[0];
        }
    }
    if (INJECT) {
#pragma unroll
        for (int j = 0; j < 2; ++j) u[j] = __ldg(a.uniforms + Prod::uidx(r0 + rr[j], ss[j], a.n_samples));
    } else {
        // sample s uses Philox counter s>>1, word pair s&1: adjacent samples of one row share a call
        const uint64_t ra = r0 + rr[0];
        const uint32_t rid0 = LONG ? rid_long : (a.row_ids ? __ldg(a.row_ids + ra) : (uint32_t)ra);
        const Philox4 q = a.rounds == 7 ? philox4x32_rk<7>(ss[0] >> 1, 0u, rid0, Prod::kTag, a.rk)
                                        : philox4x32_rk<10>(ss[0] >> 1, 0u, rid0, Prod::kTag, a.rk);
        if (PAIRED) {                        // ss[0] is even
            u[0] = u53(q.x, q.y);
            u[1] = u53(q.z, q.w);
        } else {
            u[0] = (ss[0] & 1u) ? u53(q.z, q.w) : u53(q.x, q.y);
            if (rr[1] == rr[0] && (ss[1] >> 1) == (ss[0] >> 1)) {
                u[1] = (ss[1] & 1u) ? u53(q.z, q.w) : u53(q.x, q.y);
            } else {
                const uint64_t rb = r0 + rr[1];
                const uint32_t rid1 = a.row_ids ? __ldg(a.row_ids + rb) : (uint32_t)rb;
                u[1] = uniform_1(a.seed, Prod::kTag, rid1, ss[1], a.rounds);
            }
        }
    }
    // (a lone last sample computes a throw-away duplicate in slot 1: keeps the pair code branch-free)
    if (FAST) {
        // straight-line arithmetic for both samples (the independent exp / rsqrt / rcp chains interleave); the rare
        // samples outside the branch-free domain are redone afterwards
#pragma unroll
        for (int j = 0; j < 2; ++j) e[j] = Prod::sample_fast(u[j], rows + (size_t)rr[j] * NP, a.g, f[j], s_tab);
        if (STAGE >= 1) {
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 2; ++j)
                ok &= (PAIRED || !a.prop.short_lived) ? propagate_try_fast<SHORT>(e[j], f[j], a.prop, d64v[j], s64v[j], s_tab)
                                                      : propagate_try_fast<true>(e[j], f[j], a.prop, d64v[j], s64v[j], s_tab);
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
    } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            e[j] = Prod::sample(u[j], rows + (size_t)rr[j] * NP, a.g, f[j]);
            if (STAGE >= 1) {
                propagate_sample(e[j], f[j], a.prop, d64v[j], s64v[j]);
                d32v[j] = prop_cast(d64v[j], a.prop);
                s32v[j] = prop_cast(s64v[j], a.prop);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        if (STAGE >= 2) {
            evv[j] = FAST ? decay_event_weight_coupled(e[j], d32v[j], a.ev) : decay_event_weight(e[j], d32v[j], a.ev);
            if (j == 0 || have2) acc += evv[j];
        }
        if (HIST && (j == 0 || have2)) {
            // np.histogram(E_a, bins=edges, weights=<weight of the last fused stage>)
            const double hw = STAGE >= 2 ? evv[j] : (STAGE == 1 ? (double)d32v[j] : f[j]);
            const int bin = find_bin_guess(e[j], s_edges, a.n_edges, guess);
            // s_bins: this warp's copy of the bins (or the CTA's).  Full tiles keep every lane of the warp here, so the
            // aggregated atomic-free add applies; partial tiles fall back to shared atomics
            if (!CHECK && PAIRED && a.hist_copies > 1) warp_hist_add(s_bins, bin, hw, bin >= 0 && hw != 0.0);
            else if (bin >= 0 && hw != 0.0) atomicAdd(s_bins + bin, hw);
        }
    }
    if (OUT == kOutNone) return;
    constexpr bool kAll = OUT == kOutStandard || OUT == kOutNoEvents;   // pointers known non-null (d64 / s64 known null)
    constexpr bool kEvw = OUT == kOutStandard;                            // event_w known non-null
    if (have2) {
        if (kAll || a.out_e) *reinterpret_cast<double2*>(a.out_e + i0) = make_double2(e[0], e[1]);
        if (kAll || a.out_f) *reinterpret_cast<double2*>(a.out_f + i0) = make_double2(f[0], f[1]);
        if (STAGE >= 1) {
            if (kAll || a.d32) *reinterpret_cast<float2*>(a.d32 + i0) = make_float2(d32v[0], d32v[1]);
            if (kAll || a.s32) *reinterpret_cast<float2*>(a.s32 + i0) = make_float2(s32v[0], s32v[1]);
            if (!kAll && a.d64) *reinterpret_cast<double2*>(a.d64 + i0) = make_double2(d64v[0], d64v[1]);
            if (!kAll && a.s64) *reinterpret_cast<double2*>(a.s64 + i0) = make_double2(s64v[0], s64v[1]);
        }
        if (STAGE >= 2 && (kEvw || (OUT == kOutGeneric && a.event_w))) *reinterpret_cast<double2*>(a.event_w + i0) = make_double2(evv[0], evv[1]);
    } else {
        if (kAll || a.out_e) a.out_e[i0] = e[0];
        if (kAll || a.out_f) a.out_f[i0] = f[0];
        if (STAGE >= 1) {
            if (kAll || a.d32) a.d32[i0] = d32v[0];
            if (kAll || a.s32) a.s32[i0] = s32v[0];
            if (!kAll && a.d64) a.d64[i0] = d64v[0];
            if (!kAll && a.s64) a.s64[i0] = s64v[0];
        }
        if (STAGE >= 2 && (kEvw || (OUT == kOutGeneric && a.event_w))) a.event_w[i0] = evv[0];
    }
}

// SHORT: the short-lived variant of the straight-line propagate (see propagate_try_fast); only with PAIRED && FAST.
template <class Prod, bool INJECT, int STAGE, bool FAST, bool HIST, bool PAIRED, int OUT, bool SHORT>   // STAGE 0: production; 1: + propagate; 2: + decays
__global__ void __launch_bounds__(kThreads, ALPK_MIN_BLOCKS)
chain_kernel(const ChainArgs<Prod> a) {
    using namespace alp;
    constexpr int NP = Prod::NP;
    __shared__ double s_rows[kStageRows * NP];
    __shared__ double red[32];
    __shared__ double s_tab[alp::kExpTabSize];                     // exp_fast table (fast mode)
    extern __shared__ double s_hist_mem[];           // [n_edges] edges + [n_edges-1] privatised bins (only if a.hist)
    double* s_edges = s_hist_mem;
    double* s_bins_all = s_hist_mem + a.n_edges;
    double* s_bins = s_bins_all;
    BinGuess guess{0, 0.0, 0.0};
    if (HIST) {
        for (int i = threadIdx.x; i < a.n_edges; i += blockDim.x) s_edges[i] = __ldg(a.hist_edges + i);
        for (int i = threadIdx.x; i < (a.n_edges - 1) * a.hist_copies; i += blockDim.x) s_bins_all[i] = 0.0;
        __syncthreads();
        guess = detect_spacing(s_edges, a.n_edges);
        if (a.hist_copies > 1) s_bins = s_bins_all + (threadIdx.x >> 5) * (a.n_edges - 1);
    }
    if (FAST) for (int i = threadIdx.x; i < alp::kExpTabSize; i += blockDim.x) s_tab[i] = kExpTabDev[i];
    // (all made visible by the __syncthreads of the first tile)
    const uint64_t n = a.n_rows * (uint64_t)a.n_samples;
    const uint32_t tile_elems = a.tile_elems;        // kTile, or fewer when rows are so short that a full tile would
    const uint64_t n_tiles = (n + tile_elems - 1) / tile_elems;   // touch more than kStageRows of them
    const bool long_rows = a.n_samples >= tile_elems;              // a tile then touches at most two rows
    double acc = 0.0;

    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t base = tile * tile_elems;
        const uint64_t left = n - base;
        const uint32_t lim = left < (uint64_t)tile_elems ? (uint32_t)left : tile_elems;   // samples in this tile
        const uint64_t r0 = base / a.n_samples;       // the only 64-bit division, once per tile
        const uint32_t off_last = (uint32_t)((base + lim - 1) - r0 * (uint64_t)a.n_samples);
        const int nrows = (int)(long_rows ? (off_last >= a.n_samples ? 2u : 1u) : off_last / a.n_samples + 1u);   // <= kStageRows
        __syncthreads();                             // previous tile finished reading s_rows
        stage_rows<NP>(s_rows, a.table, r0, nrows);
        __syncthreads();
        const uint32_t off0 = (uint32_t)(base - r0 * (uint64_t)a.n_samples);   // offset of the tile in its first row

        if (PAIRED && !INJECT && lim == (uint32_t)kTile && long_rows) {
            // full tile of long rows: no bounds tests, row bookkeeping hoisted to the tile
            TileRows tr;
            tr.split = a.n_samples - off0;             // samples of this tile that belong to row r0 (>= kTile: all of them)
            tr.rid0 = a.row_ids ? __ldg(a.row_ids + r0) : (uint32_t)r0;
            tr.rid1 = (tr.split < (uint32_t)kTile) ? (a.row_ids ? __ldg(a.row_ids + r0 + 1) : (uint32_t)(r0 + 1)) : tr.rid0;
            const uint32_t local0 = 2 * threadIdx.x;
#pragma unroll kInnerUnroll
            for (int k = 0; k < kUnroll; ++k)
                chain_pair<Prod, INJECT, STAGE, FAST, HIST, PAIRED, OUT, SHORT, false, true>(
                    a, s_rows, s_tab, s_edges, s_bins, r0, off0, base, local0 + (uint32_t)k * (2 * kThreads), lim, true, acc, tr, guess);
        } else if (PAIRED && lim == (uint32_t)kTile) {       // full tile: no bounds tests in the loop
#pragma unroll kInnerUnroll
            for (int k = 0; k < kUnroll; ++k)
                chain_pair<Prod, INJECT, STAGE, FAST, HIST, PAIRED, OUT, SHORT, false>(
                    a, s_rows, s_tab, s_edges, s_bins, r0, off0, base, 2 * ((uint32_t)k * kThreads + threadIdx.x), lim, long_rows, acc,
                    TileRows{0u, 0u, 0u}, guess);
        } else {
#pragma unroll 1
            for (int k = 0; k < kUnroll; ++k) {
                const uint32_t local = 2 * ((uint32_t)k * kThreads + threadIdx.x);
                if (local >= lim) break;
                chain_pair<Prod, INJECT, STAGE, FAST, HIST, PAIRED, OUT, SHORT, true>(
                    a, s_rows, s_tab, s_edges, s_bins, r0, off0, base, local, lim, long_rows, acc, TileRows{0u, 0u, 0u}, guess);
            }
        }
    }
    if (STAGE >= 2) {
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) a.partials[blockIdx.x] = acc;
    }
    if (HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < a.n_edges - 1; i += blockDim.x) {
            double v = 0.0;
            for (int c = 0; c < a.hist_copies; ++c) v += s_bins_all[c * (a.n_edges - 1) + i];
            if (v != 0.0) atomicAdd(a.hist + i, v);
        }
    }
}

alp::PropConst to_prop(const alpk_prop_t* p);
alp::EventConst to_event(const alpk_event_t* e);
int finalize_sum(const double* partials, int n, double* out, cudaStream_t st);

// Row-aligned warp-unit kernel (chain_rows.cuh); returns -1 when the configuration is not one it covers.
template <class Prod>
int launch_chain_rows(const char* what, ChainArgs<Prod>& a, int stage, bool standard, bool noevents, bool none,
                      double* rate, double* scratch, cudaStream_t st);

// Host-side launcher shared by every production channel.
template <class Prod>
int launch_chain(const char* what, const double* row_table, uint64_t n_rows, uint32_t n_samples, const uint32_t* row_ids,
                 const typename Prod::Globals& g, int fast_mode, int rng_rounds, const double* uniforms, uint64_t seed, double* out_energy,
                 double* out_flux, const alpk_prop_t* h_prop, float* decay_w32, float* scatter_w32, double* decay_w64,
                 double* scatter_w64, const alpk_event_t* h_ev, double* event_w, double* rate, double* scratch,
                 const double* hist_edges, int n_edges, double* hist, void* stream) {
    if (h_ev && (!h_prop || !rate || !scratch)) { set_error("%s: events need prop, rate and scratch", what); return 2; }
    if (n_samples == 0 || n_samples > 0x7fffffffu - kTile) { set_error("%s: n_samples out of range", what); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const uint64_t n = n_rows * (uint64_t)n_samples;
    if (n == 0) {
        if (h_ev) { cudaMemsetAsync(rate, 0, sizeof(double), st); }
        return 0;
    }
    if (!row_table) { set_error("%s: null row table", what); return 2; }
    // pairs of samples are stored as 16-byte (f64) / 8-byte (f32) vectors: reject misaligned outputs here rather than fault
    // asynchronously in the kernel
    if ((((uintptr_t)out_energy | (uintptr_t)out_flux | (uintptr_t)decay_w64 | (uintptr_t)scatter_w64 | (uintptr_t)event_w) & 15u) ||
        (((uintptr_t)decay_w32 | (uintptr_t)scatter_w32) & 7u)) {
        set_error("%s: per-sample output arrays must be 16-byte aligned (float32 arrays: 8-byte)", what);
        return 2;
    }
    ChainArgs<Prod> a;
    memset(&a, 0, sizeof(a));
    a.table = row_table; a.n_rows = n_rows; a.n_samples = n_samples; a.row_ids = row_ids;
    a.g = g; a.uniforms = uniforms; a.seed = seed; a.rk = alp::philox_keys(seed);
    if (rng_rounds != 0 && rng_rounds != 7 && rng_rounds != 10) { set_error("%s: rng_rounds must be 10 (or 0) or 7", what); return 2; }
    a.rounds = rng_rounds == 7 ? 7 : 10;
    a.out_e = out_energy; a.out_f = out_flux;
    const int stage = h_ev ? 2 : (h_prop ? 1 : 0);
    if (h_prop) { a.prop = to_prop(h_prop); a.d32 = decay_w32; a.s32 = scatter_w32; a.d64 = decay_w64; a.s64 = scatter_w64; }
    if (h_ev) { a.ev = to_event(h_ev); a.event_w = event_w; a.partials = scratch; }
    size_t smem = 0;
    if (hist) {
        if (!hist_edges || n_edges < 2) { set_error("%s: histogram needs >= 2 edges", what); return 2; }
        const size_t smem8 = (size_t)(n_edges + (kThreads / 32) * (n_edges - 1)) * sizeof(double);
        a.hist_copies = smem8 <= 40 * 1024 ? kThreads / 32 : 1;
        smem = a.hist_copies > 1 ? smem8 : (size_t)(2 * n_edges - 1) * sizeof(double);
        if (smem > 40 * 1024) { set_error("%s: fused histogram supports at most %d edges (use alpk_hist)", what, (40 * 1024 / 8 + 1) / 2); return 2; }
        a.hist_edges = hist_edges; a.n_edges = n_edges; a.hist = hist;
    }
    // rows per tile <= tile/n_samples + 2 must fit the shared-memory stage
    uint64_t te = (uint64_t)(kStageRows - 2) * n_samples;
    a.tile_elems = (uint32_t)(te >= (uint64_t)kTile ? (uint64_t)kTile : (te < 2 ? 2 : (te & ~1ull)));
    const int grid = grid_for(n, (int)a.tile_elems, ALPK_GRID_PER_SM);
    const bool inj = uniforms != nullptr;
    const bool fast = fast_mode != 0;
    // the PAIRED / output-set specialisations exist for the production configuration (fast arithmetic, free-running RNG);
    // the row kernel takes it, with or without the fused histogram, whenever its geometry applies
    const bool paired = (n_samples % 2u) == 0u;
    const bool none = !out_energy && !out_flux && !decay_w32 && !scatter_w32 && !decay_w64 && !scatter_w64 && !event_w;
    const bool standard = out_energy && out_flux && !decay_w64 && !scatter_w64 && (stage < 1 || (decay_w32 && scatter_w32)) &&
                          (stage < 2 || event_w);
    const bool noevents = stage == 2 && out_energy && out_flux && decay_w32 && scatter_w32 && !decay_w64 && !scatter_w64 && !event_w;
    if (fast && paired && !inj) {
        const int rc = launch_chain_rows<Prod>(what, a, stage, standard, noevents, none, rate, scratch, st);
        if (rc >= 0) return rc;
    }
    // (histogram launches: static shared memory -- row stage + 16 KB exp table -- plus the bins can exceed the 48 KB a kernel
    // gets without opting in)
#define ALPK_CHAIN_LAUNCH(I, S, F, H, P, O, SH) do {                                                                   \
        auto kfn = chain_kernel<Prod, I, S, F, H, P, O, SH>;                                                          \
        if (smem > 16 * 1024) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        kfn<<<grid, kThreads, smem, st>>>(a); } while (0)
#define ALPK_CHAIN_PAIRED(S, O) { if (S >= 1 && a.prop.short_lived) ALPK_CHAIN_LAUNCH(false, S, true, false, true, O, true); \
                                  else ALPK_CHAIN_LAUNCH(false, S, true, false, true, O, false); }
#define ALPK_CHAIN_CASE(S)                                                                                   \
    if (hist) {   /* histogram variants only for the free-running RNG */                                      \
        if (inj) { set_error("%s: fused histogram is not available with injected uniforms", what); return 2; } \
        if (fast && paired) ALPK_CHAIN_LAUNCH(false, S, true, true, true, kOutGeneric, false);                   \
        else if (fast) ALPK_CHAIN_LAUNCH(false, S, true, true, false, kOutGeneric, false);                       \
        else ALPK_CHAIN_LAUNCH(false, S, false, true, false, kOutGeneric, false);                                \
    } else if (inj) { if (fast) ALPK_CHAIN_LAUNCH(true, S, true, false, false, kOutGeneric, false); else ALPK_CHAIN_LAUNCH(true, S, false, false, false, kOutGeneric, false); } \
    else if (fast && paired) {                                                                                 \
        if (standard) ALPK_CHAIN_PAIRED(S, kOutStandard)                                                       \
        else if (S == 2 && noevents) ALPK_CHAIN_PAIRED(S, kOutNoEvents)                                        \
        else if (none) ALPK_CHAIN_PAIRED(S, kOutNone)                                                          \
        else ALPK_CHAIN_PAIRED(S, kOutGeneric)                                                                 \
    } else if (fast) ALPK_CHAIN_LAUNCH(false, S, true, false, false, kOutGeneric, false);                      \
    else ALPK_CHAIN_LAUNCH(false, S, false, false, false, kOutGeneric, false);
    if (stage == 0) { ALPK_CHAIN_CASE(0) } else if (stage == 1) { ALPK_CHAIN_CASE(1) } else { ALPK_CHAIN_CASE(2) }
#undef ALPK_CHAIN_LAUNCH
#undef ALPK_CHAIN_PAIRED
#undef ALPK_CHAIN_CASE
    if (check_launch(what)) return 1;
    if (stage == 2) return finalize_sum(scratch, grid, rate, st);
    return 0;
}

}  // namespace alpk

#include "chain_rows.cuh"
