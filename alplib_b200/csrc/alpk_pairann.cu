// libalpk: associated production e+ e- -> a gamma (FluxPairAnnihilationIsotropic, fluxes.py:469-549) -- its own
// translation unit so the chain-kernel instantiations of the e+- channels compile in parallel.
#include <string.h>

#include "../../include/alpk.h"
#include "alp_math2.cuh"
#include "chain_kernel.cuh"

using namespace alp;

namespace alpk {

// ---------------------------------------------------------------------------------------------------------------
// associated production
// ---------------------------------------------------------------------------------------------------------------
struct PairProd {
    static constexpr int NP = PA_NP;
    static constexpr uint32_t kTag = kTagPairAnn;
    using Globals = PairGlobals;
    // injected layout [row][2][n]: phi block then theta block; the kernel needs the theta draw
    __device__ static __forceinline__ uint64_t uidx(uint64_t row, uint32_t s, uint32_t n) { return (row * 2 + 1) * n + s; }
    __device__ static __forceinline__ double sample(double u, const double* r, const Globals& g, double& flux) {
        return pairann_sample(u, r, g, flux);
    }
    __device__ static __forceinline__ double sample_fast(double u, const double* r, const Globals& g, double& flux, const double* exp_tab) {
        return pairann_sample_fast(u, r, g, flux);
    }
};

static PairGlobals to_pair(const alpk_pairann_t* p) {
    PairGlobals g;
    g.ma = p->ma; g.ma2 = p->ma2; g.me2 = p->me2; g.me4 = p->me4; g.pref = p->pref; g.wconst = p->wconst;
    g.z = p->z; g.ge2 = p->ge2; g.ntad = p->ntad; g.hbarc2 = p->hbarc2; g.n_samples = p->n_samples;
    return g;
}

__global__ void __launch_bounds__(kThreads)
pairann_rows_kernel(const double* __restrict__ ep, const double* __restrict__ wp, uint64_t n, PairGlobals g,
                    double* __restrict__ table) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double row[PA_NP];
        pairann_row_params(__ldg(ep + i), __ldg(wp + i), g, row);
#pragma unroll
        for (int k = 0; k < PA_NP; ++k) table[i * PA_NP + k] = row[k];
    }
}

}  // namespace alpk

using namespace alpk;

ALPK_EXPORT int alpk_pairann_rows(const double* e_pos, const double* w_pos, uint64_t n_rows, const alpk_pairann_t* h_par,
                                  double* row_table, void* stream) {
    static_assert(PA_NP == ALPK_PAIRANN_NP, "header/table layout mismatch");
    if (!h_par || !row_table) { set_error("alpk_pairann_rows: null argument"); return 2; }
    if (n_rows == 0) return 0;
    if (!e_pos || !w_pos) { set_error("alpk_pairann_rows: null input"); return 2; }
    pairann_rows_kernel<<<grid_for(n_rows, kThreads, 4), kThreads, 0, (cudaStream_t)stream>>>(e_pos, w_pos, n_rows,
                                                                                         to_pair(h_par), row_table);
    return check_launch("alpk_pairann_rows");
}

ALPK_EXPORT int alpk_pairann_samples(const double* row_table, uint64_t n_rows, uint32_t n_samples, const uint32_t* row_ids,
                                     const alpk_pairann_t* h_par, const double* uniforms, uint64_t seed,
                                     double* out_energy, double* out_flux,
                                     const alpk_prop_t* h_prop, float* decay_w32, float* scatter_w32,
                                     double* decay_w64, double* scatter_w64,
                                     const alpk_event_t* h_ev, double* event_w, double* rate, double* scratch,
                                     const double* hist_edges, int n_edges, double* hist, void* stream) {
    if (!h_par) { set_error("alpk_pairann_samples: null argument"); return 2; }
    if (h_prop && (h_prop->fast != 0) != (h_par->fast != 0)) { set_error("alpk_pairann_samples: production and propagate precision modes differ"); return 2; }
    return launch_chain<PairProd>("alpk_pairann_samples", row_table, n_rows, n_samples, row_ids, to_pair(h_par),
                                  h_par->fast ? 1 : 0, h_par->rng_rounds, uniforms,
                                  seed, out_energy, out_flux, h_prop, decay_w32, scatter_w32, decay_w64, scatter_w64,
                                  h_ev, event_w, rate, scratch, hist_edges, n_edges, hist, stream);
}
