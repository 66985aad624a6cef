// libalpk: the e+- production channels -- bremsstrahlung (track-length attenuated), resonance, associated production.
#include <string.h>

#include "../../include/alpk.h"
#include "alp_math2.cuh"
#include "chain_kernel.cuh"

using namespace alp;

namespace alpk {

// ---------------------------------------------------------------------------------------------------------------
// bremsstrahlung
// ---------------------------------------------------------------------------------------------------------------
struct BremProd {
    static constexpr int NP = BR_NP;
    static constexpr uint32_t kTag = kTagBremEa;
    using Globals = BremGlobals;
    __device__ static __forceinline__ uint64_t uidx(uint64_t row, uint32_t s, uint32_t n) { return row * n + s; }
    __device__ static __forceinline__ double sample(double u, const double* b, const Globals& g, double& flux) {
        return brem_sample(u, b, g, flux);
    }
    __device__ static __forceinline__ double sample_fast(double u, const double* b, const Globals& g, double& flux, const double* exp_tab) {
        return brem_sample_fast(u, b, g, flux, exp_tab);
    }
};

static BremGlobals to_brem(const alpk_brem_t* p) {
    BremGlobals g;
    g.ma = p->ma; g.ma2 = p->ma2; g.log10_ma = p->log10_ma; g.ep_min = p->ep_min; g.nt = p->nt; g.pref_num = p->pref_num;
    g.zl = p->zl; g.zz = p->zz; g.pref_vec = p->pref_vec; g.ln10 = p->ln10; g.n_samples = p->n_samples; g.t_arg = p->t_arg;
    g.vector = p->vector; g.use_track_length = p->use_track_length;
    return g;
}

// one thread per electron row: scalar draw, dN/dE lookups, Gauss-Legendre track-length integral, sampling bounds
__global__ void __launch_bounds__(128)
brem_rows_kernel(const double* __restrict__ e0, const double* __restrict__ w0, uint64_t n, const uint32_t* __restrict__ row_ids,
                 const double* __restrict__ u_rows, uint64_t seed, const double* __restrict__ ele_e,
                 const double* __restrict__ ele_f, int n_ele, const double* __restrict__ pos_e,
                 const double* __restrict__ pos_f, int n_pos, const double* __restrict__ gl_x,
                 const double* __restrict__ gl_w, BremGlobals g, double* __restrict__ table, uint8_t* __restrict__ emit) {
    __shared__ double s_x[kGL], s_w[kGL];
    if (threadIdx.x < kGL) { s_x[threadIdx.x] = gl_x[threadIdx.x]; s_w[threadIdx.x] = gl_w[threadIdx.x]; }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = __ldg(e0 + i);
        double u = 0.0, fl = 0.0;
        if (g.use_track_length) {
            const uint32_t rid = row_ids ? __ldg(row_ids + i) : (uint32_t)i;
            u = u_rows ? __ldg(u_rows + i) : uniform_1(seed, kTagBremRow, rid, 0);
            fl = interp(e, ele_e, ele_f, n_ele, 0.0, 0.0) + interp(e, pos_e, pos_f, n_pos, 0.0, 0.0);   // fluxes.py:316
        }
        double row[BR_NP];
        brem_row_params(e, __ldg(w0 + i), u, fl, g, s_x, s_w, row);
#pragma unroll
        for (int k = 0; k < BR_NP; ++k) table[i * BR_NP + k] = row[k];
        emit[i] = row[BR_EMIT] != 0.0 ? 1 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// resonance: block-reduced Monte-Carlo integral
// ---------------------------------------------------------------------------------------------------------------
template <bool INJECT, bool FAST>
__global__ void __launch_bounds__(kThreads)
resonance_kernel(const double* __restrict__ pos_e, const double* __restrict__ pos_f, int n_pos, double eres, double emax,
                 double max_t, uint64_t n, const double* __restrict__ uniforms, uint64_t seed, double* __restrict__ partials) {
    __shared__ double red[32];
    __shared__ double s_tab[alp::kExpTabSize];
    double inv_dx = 0.0;
    if (FAST) {
        for (int i = threadIdx.x; i < alp::kExpTabSize; i += blockDim.x) s_tab[i] = kExpTabDev[i];
        // uniformly spaced positron table (the usual linspace): index from the spacing instead of a binary search
        const BinGuess g = detect_spacing(pos_e, n_pos);      // (contains the CTA-wide barriers that also publish s_tab)
        inv_dx = g.mode == 1 ? g.scale : 0.0;
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double u1, u2;
        if (INJECT) { u1 = __ldg(uniforms + i); u2 = __ldg(uniforms + n + i); }
        else uniform_2(seed, kTagResonance, 0u, i, u1, u2);
        acc += FAST ? resonance_term_fast(u1, u2, eres, emax, max_t, pos_e, pos_f, n_pos, inv_dx, s_tab)
                    : resonance_term(u1, u2, eres, emax, max_t, pos_e, pos_f, n_pos);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

}  // namespace alpk

using namespace alpk;

ALPK_EXPORT int alpk_brem_rows(const double* e0, const double* w0, uint64_t n_rows, const uint32_t* row_ids,
                               const double* u_rows, uint64_t seed,
                               const double* ele_e, const double* ele_f, int n_ele, const double* pos_e,
                               const double* pos_f, int n_pos, const double* gl_x, const double* gl_w,
                               const alpk_brem_t* h_par, double* row_table, uint8_t* emit, void* stream) {
    static_assert(BR_NP == ALPK_BREM_NP && kGL == ALPK_GL_POINTS, "header/table layout mismatch");
    if (!h_par || !row_table || !emit || !gl_x || !gl_w) { set_error("alpk_brem_rows: null argument"); return 2; }
    if (h_par->use_track_length && (!ele_e || !ele_f || !pos_e || !pos_f || n_ele < 1 || n_pos < 1)) {
        set_error("alpk_brem_rows: track-length mode needs the electron and positron dN/dE tables"); return 2;
    }
    if (n_rows == 0) return 0;
    if (!e0 || !w0) { set_error("alpk_brem_rows: null input"); return 2; }
    brem_rows_kernel<<<grid_for(n_rows, 128, 8), 128, 0, (cudaStream_t)stream>>>(
        e0, w0, n_rows, row_ids, u_rows, seed, ele_e, ele_f, n_ele, pos_e, pos_f, n_pos, gl_x, gl_w, to_brem(h_par),
        row_table, emit);
    return check_launch("alpk_brem_rows");
}

ALPK_EXPORT int alpk_brem_samples(const double* row_table, uint64_t n_rows, uint32_t n_samples, const uint32_t* row_ids,
                                  const alpk_brem_t* h_par, const double* uniforms, uint64_t seed,
                                  double* out_energy, double* out_flux,
                                  const alpk_prop_t* h_prop, float* decay_w32, float* scatter_w32,
                                  double* decay_w64, double* scatter_w64,
                                  const alpk_event_t* h_ev, double* event_w, double* rate, double* scratch,
                                     const double* hist_edges, int n_edges, double* hist, void* stream) {
    if (!h_par) { set_error("alpk_brem_samples: null argument"); return 2; }
    if (h_prop && (h_prop->fast != 0) != (h_par->fast != 0)) { set_error("alpk_brem_samples: production and propagate precision modes differ"); return 2; }
    return launch_chain<BremProd>("alpk_brem_samples", row_table, n_rows, n_samples, row_ids, to_brem(h_par),
                                  h_par->fast ? 1 : 0, h_par->rng_rounds, uniforms,
                                  seed, out_energy, out_flux, h_prop, decay_w32, scatter_w32, decay_w64, scatter_w64,
                                  h_ev, event_w, rate, scratch, hist_edges, n_edges, hist, stream);
}

ALPK_EXPORT int alpk_resonance_sum(const double* pos_e, const double* pos_f, int n_pos, double e_res, double e_max,
                                   double max_t, uint64_t n_samples, const double* uniforms, uint64_t seed, int fast,
                                   double* out_sum, double* scratch, void* stream) {
    if (!pos_e || !pos_f || n_pos < 1 || !out_sum || !scratch) { set_error("alpk_resonance_sum: bad argument"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n_samples == 0) { cudaMemsetAsync(out_sum, 0, sizeof(double), st); return 0; }
    const int grid = grid_for(n_samples, kThreads * 4, 8);
#define ALPK_RES(I, F) resonance_kernel<I, F><<<grid, kThreads, 0, st>>>(pos_e, pos_f, n_pos, e_res, e_max, max_t, n_samples, uniforms, seed, scratch)
    if (uniforms) { if (fast) ALPK_RES(true, true); else ALPK_RES(true, false); }
    else { if (fast) ALPK_RES(false, true); else ALPK_RES(false, false); }
#undef ALPK_RES
    if (check_launch("alpk_resonance_sum")) return 1;
    return finalize_sum(scratch, grid, out_sum, st);
}
