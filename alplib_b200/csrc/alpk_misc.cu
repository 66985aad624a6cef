// libalpk: small flux sources beyond the photon / e+- channels -- nuclear de-excitation lines (fluxes.py:554-600) and
// pi0 -> gamma a two-body decay in flight (fluxes.py:926-991).  One thread per input row.
#include "../../include/alpk.h"
#include "alp_math2.cuh"
#include "common.cuh"

using namespace alp;

namespace alpk {

// fluxes.py:571-577 (br) and :587-590 (simulate); rows already filtered by the host (E > ma)
__global__ void __launch_bounds__(kThreads)
nuclear_kernel(const double* __restrict__ rates4, uint64_t n, double ma2, double c0, double expo, double gann0,
               double gann1, double* __restrict__ out_e, double* __restrict__ out_f) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = __ldg(rates4 + 4 * i), rate = __ldg(rates4 + 4 * i + 1);
        const double beta = __ldg(rates4 + 4 * i + 2), eta = __ldg(rates4 + 4 * i + 3);
        const double mu0 = 0.88, mu1 = 4.71;
        const double kin = pow(sqrt(e * e - ma2) / e, expo);
        const double ratio = (gann0 * beta + gann1) / ((mu0 - 0.5) * beta + (mu1 - eta));
        const double br = (c0 * kin) * (ratio * ratio);
        out_e[i] = e;
        out_f[i] = rate * br;
    }
}

// fluxes.py:955-968: one isotropic decay per pi0 momentum, boosted along z; keep[i] = passes the energy/angle cuts
__global__ void __launch_bounds__(kThreads)
pi0_flux_kernel(const double* __restrict__ p, uint64_t n, const double* __restrict__ uniforms, uint64_t seed,
                double mm2, double ma2, double p_cm, double e1_cm, double energy_cut, double angle_cut, double wgt,
                double* __restrict__ out_e, double* __restrict__ out_f, uint8_t* __restrict__ keep) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double u = uniforms ? __ldg(uniforms + i) : uniform_1(seed, 8u /* pi0 stream */, 0u, i);
        const double c = -1 + (1 - (-1)) * u;                                    // np.random.uniform(-1, 1)
        const double pm = __ldg(p + i);
        const double beta = pm / sqrt(pm * pm + mm2);                             // :959
        const double boost = pow(1 - beta * beta, -0.5);                          // :960
        const double e_lab = boost * (e1_cm + (beta * p_cm) * c);                 // :961
        const double pz_lab = boost * (p_cm * c + beta * e1_cm);                  // :962
        const double angle = acos(pz_lab / sqrt(e_lab * e_lab - ma2));            // :963
        out_e[i] = e_lab;
        out_f[i] = wgt;
        keep[i] = (!(e_lab < energy_cut) && !(angle > angle_cut)) ? 1 : 0;        // :964-967
    }
}

// fluxes.py:1090-1134: thread = sample; bins are the positron-table bin centres above threshold (host-filtered)
__global__ void __launch_bounds__(kThreads)
pairgamma_kernel(const double* __restrict__ centers, const double* __restrict__ widths, uint64_t n_bins, uint32_t n_samples,
                 const uint32_t* __restrict__ bin_ids, const double* __restrict__ pos_e, const double* __restrict__ pos_f,
                 int n_pos, PairGammaGlobals g, const double* __restrict__ u_bins, const double* __restrict__ u_alp,
                 uint64_t seed, double* __restrict__ out_e, double* __restrict__ out_f) {
    const uint64_t n = n_bins * (uint64_t)n_samples;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t b = i / n_samples;
        const uint32_t sidx = (uint32_t)(i - b * n_samples);
        const double e_lab = __ldg(centers + b);
        const double fc = interp(e_lab, pos_e, pos_f, n_pos, 0.0, 0.0);
        double u1, u2, u3;
        if (u_bins) {
            u1 = __ldg(u_bins + (b * 2) * n_samples + sidx);        // [bin][2][n]: depths then smeared energies
            u2 = __ldg(u_bins + (b * 2 + 1) * n_samples + sidx);
            u3 = __ldg(u_alp + i);                                  // then one draw per sample, in sample order
        } else {
            const uint32_t rid = bin_ids ? __ldg(bin_ids + b) : (uint32_t)b;
            uniform_2(seed, 9u, rid, sidx, u1, u2);
            u3 = uniform_1(seed, 10u, rid, sidx);
        }
        double f;
        out_e[i] = pairgamma_sample(u1, u2, u3, e_lab, __ldg(widths + b), fc, g, f);
        out_f[i] = f;
    }
}

// fluxes.py:1202-1224 FluxAxionMesonMixing.simulate: one ALP per meson 4-vector [E, px, py, pz] that is above the mass and
// inside the detector's half opening angle; every kept sample carries the same rate
__global__ void __launch_bounds__(kThreads)
meson_mixing_kernel(const double* __restrict__ p4, uint64_t n, double ma, double det_sa, double rate,
                    double* __restrict__ out_e, double* __restrict__ out_angle, double* __restrict__ out_f,
                    uint8_t* __restrict__ keep) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = __ldg(p4 + 4 * i), px = __ldg(p4 + 4 * i + 1), py = __ldg(p4 + 4 * i + 2), pz = __ldg(p4 + 4 * i + 3);
        const double angle = acos(pz / sqrt((px * px + py * py) + pz * pz));          // :1214
        out_e[i] = e; out_angle[i] = angle; out_f[i] = rate;
        keep[i] = (!(e < ma) && !(angle > det_sa)) ? 1 : 0;                           // :1211-1217
    }
}

__global__ void __launch_bounds__(kThreads)
scale_kernel(const double* __restrict__ x, uint64_t n, double factor, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = factor * __ldg(x + i);
}

}  // namespace alpk

using namespace alpk;

ALPK_EXPORT int alpk_pairgamma(const double* centers, const double* widths, uint64_t n_bins, uint32_t n_samples,
                               const uint32_t* bin_ids, const double* pos_e, const double* pos_f, int n_pos,
                               const alpk_pairgamma_t* h_par, const double* u_bins, const double* u_alp, uint64_t seed,
                               double* out_energy, double* out_flux, void* stream) {
    if (!h_par || !out_energy || !out_flux) { set_error("alpk_pairgamma: null argument"); return 2; }
    if (n_bins == 0 || n_samples == 0) return 0;
    if (!centers || !widths || !pos_e || !pos_f || n_pos < 1) { set_error("alpk_pairgamma: null input"); return 2; }
    if ((u_bins == nullptr) != (u_alp == nullptr)) { set_error("alpk_pairgamma: inject both uniform blocks or none"); return 2; }
    PairGammaGlobals g;
    g.ma = h_par->ma; g.ma2 = h_par->ma2; g.ma4 = h_par->ma4; g.me2 = h_par->me2; g.me4 = h_par->me4; g.ep_min = h_par->ep_min;
    g.nt = h_par->nt; g.zc = h_par->zc; g.n_samples = h_par->n_samples; g.max_t = h_par->max_t; g.fast = h_par->fast;
    pairgamma_kernel<<<grid_for(n_bins * (uint64_t)n_samples, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
        centers, widths, n_bins, n_samples, bin_ids, pos_e, pos_f, n_pos, g, u_bins, u_alp, seed, out_energy, out_flux);
    return check_launch("alpk_pairgamma");
}

ALPK_EXPORT int alpk_meson_mixing(const double* p4, uint64_t n, double ma, double det_sa, double rate, double* out_energy,
                                  double* out_angle, double* out_flux, uint8_t* keep, void* stream) {
    if (n == 0) return 0;
    if (!p4 || !out_energy || !out_angle || !out_flux || !keep) { set_error("alpk_meson_mixing: null argument"); return 2; }
    meson_mixing_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(p4, n, ma, det_sa, rate, out_energy,
                                                                                        out_angle, out_flux, keep);
    return check_launch("alpk_meson_mixing");
}

// planes[ncomp][n] (what the decay / scatter kernels write: coalesced per component) -> rows[n][ncomp] (the layout of the
// reference's per-event `pmu` arrays), so a host copy of all events is ONE contiguous transfer instead of a strided gather.
template <int NC>
__global__ void __launch_bounds__(kThreads)
planes_to_rows_kernel(const double* __restrict__ planes, uint64_t n, double* __restrict__ rows) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) v[c] = __ldg(planes + (uint64_t)c * n + i);
        if (NC == 4) {
            reinterpret_cast<double2*>(rows + 4 * i)[0] = make_double2(v[0], v[1]);
            reinterpret_cast<double2*>(rows + 4 * i)[1] = make_double2(v[2], v[3]);
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) rows[(uint64_t)NC * i + c] = v[c];
        }
    }
}

ALPK_EXPORT int alpk_planes_to_rows(const double* planes, uint64_t n, int ncomp, double* rows, void* stream) {
    if (n == 0) return 0;
    if (!planes || !rows) { set_error("alpk_planes_to_rows: null argument"); return 2; }
    if (ncomp != 3 && ncomp != 4) { set_error("alpk_planes_to_rows: ncomp must be 3 or 4"); return 2; }
    if (ncomp == 4 && ((uintptr_t)rows & 15)) { set_error("alpk_planes_to_rows: rows must be 16-byte aligned"); return 2; }
    const int grid = grid_for(n, kThreads, 8);
    if (ncomp == 4) planes_to_rows_kernel<4><<<grid, kThreads, 0, (cudaStream_t)stream>>>(planes, n, rows);
    else planes_to_rows_kernel<3><<<grid, kThreads, 0, (cudaStream_t)stream>>>(planes, n, rows);
    return check_launch("alpk_planes_to_rows");
}

ALPK_EXPORT int alpk_scale(const double* x, uint64_t n, double factor, double* out, void* stream) {
    if (n == 0) return 0;
    if (!x || !out) { set_error("alpk_scale: null argument"); return 2; }
    scale_kernel<<<grid_for(n, kThreads * 4, 8), kThreads, 0, (cudaStream_t)stream>>>(x, n, factor, out);
    return check_launch("alpk_scale");
}

ALPK_EXPORT int alpk_nuclear(const double* rates4, uint64_t n_rows, double ma2, double c0, double expo, double gann0,
                             double gann1, double* out_energy, double* out_flux, void* stream) {
    if (n_rows == 0) return 0;
    if (!rates4 || !out_energy || !out_flux) { set_error("alpk_nuclear: null argument"); return 2; }
    nuclear_kernel<<<grid_for(n_rows, kThreads, 4), kThreads, 0, (cudaStream_t)stream>>>(rates4, n_rows, ma2, c0, expo, gann0,
                                                                                       gann1, out_energy, out_flux);
    return check_launch("alpk_nuclear");
}

ALPK_EXPORT int alpk_pi0_flux(const double* momenta, uint64_t n, const double* uniforms, uint64_t seed, double mm2,
                              double ma2, double p_cm, double e1_cm, double energy_cut, double angle_cut, double weight,
                              double* out_energy, double* out_flux, uint8_t* keep, void* stream) {
    if (n == 0) return 0;
    if (!momenta || !out_energy || !out_flux || !keep) { set_error("alpk_pi0_flux: null argument"); return 2; }
    pi0_flux_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
        momenta, n, uniforms, seed, mm2, ma2, p_cm, e1_cm, energy_cut, angle_cut, weight, out_energy, out_flux, keep);
    return check_launch("alpk_pi0_flux");
}
