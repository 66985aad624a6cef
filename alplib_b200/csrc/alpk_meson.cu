// libalpk: charged-meson 3-body decay fluxes  M -> l nu X  (fluxes.py:605-922): FluxChargedMeson3BodyIsotropic and
// FluxChargedMeson3BodyDecay.  One (meson, lepton) pair = one row; thread = sample; every sample evaluates the Dalitz
// integral dGamma/dE_a with the 21-point Gauss-Kronrod rule QUADPACK uses for the reference's scipy.integrate.quad
// (alp_math3.cuh), i.e. >= 21 matrix-element evaluations per sample: FP64-issue bound, no memory traffic to speak of.
#include "../../include/alpk.h"
#include "alp_math3.cuh"
#include "common.cuh"

using namespace alp;

namespace alpk {

constexpr uint32_t kTagMeson3Pos = 11, kTagMeson3EC = 12, kTagMeson3Phi = 13;

static Meson3Const to_meson3(const alpk_meson3_t* p) {
    Meson3Const K;
    K.mm = p->mm; K.ma = p->ma; K.fM = p->fM; K.ckm = p->ckm; K.total_width = p->total_width; K.c0 = p->c0;
    K.coupling = p->coupling; K.a = p->a; K.b = p->b; K.d = p->d; K.pref = p->pref; K.n_samples = p->n_samples;
    K.model = p->model;
    return K;
}

// rows: one per (meson, allowed lepton).  mode 0: isotropic (mesons[n][2] = p, weight); mode 1: decay in flight
// (mesons[n][3] = p, theta, weight; decay position x = -log(1-u)*scale, scale = p*PION_LIFETIME*0.01*C_LIGHT/M_PI, :740)
__global__ void __launch_bounds__(kThreads)
meson3_rows_kernel(int mode, const double* __restrict__ mesons, const uint32_t* __restrict__ row_meson,
                   const double* __restrict__ row_ml, uint64_t n_rows, Meson3Const K, double det_dist, double det_length,
                   const double* __restrict__ u_pos, uint64_t seed, uint32_t meson_offset, double* __restrict__ table) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const uint32_t im = __ldg(row_meson + r);
        const double ml = __ldg(row_ml + r);
        double row[MR_NP];
        if (mode == 0) {
            meson3_row_isotropic(__ldg(mesons + 2 * (size_t)im), __ldg(mesons + 2 * (size_t)im + 1), ml, K, row);
        } else {
            const double p = __ldg(mesons + 3 * (size_t)im);
            const double u = u_pos ? __ldg(u_pos + im) : uniform_1(seed, kTagMeson3Pos, 0u, (uint64_t)im + meson_offset);
            const double scale = (((p * 2.6e-8) * 0.01) * kCLight) / 139.57;      // PION_LIFETIME, M_PI (constants.py:108, 41)
            const double x = (-log(1.0 - u)) * scale + 0.0;                       // expon.rvs: standard_exponential*scale + loc
            meson3_row_decay(p, __ldg(mesons + 3 * (size_t)im + 1), __ldg(mesons + 3 * (size_t)im + 2), ml, x, det_dist,
                             det_length, K, row);
        }
#pragma unroll
        for (int k = 0; k < MR_NP; ++k) table[r * MR_NP + k] = row[k];
    }
}

// samples: thread = (row, sample).  Injected uniforms: block_of_row[r] >= 0 is the index of the row's block in
// `uniforms`, laid out [block][2 or 3][n_samples] (energies, cosines[, azimuths] -- the reference's draw order).
__global__ void __launch_bounds__(kThreads, 2)
meson3_samples_kernel(int mode, const double* __restrict__ table, uint64_t n_rows, uint32_t n_samples, Meson3Const K,
                      double energy_cut, int cut_on_solid_angle, const double* __restrict__ uniforms,
                      const int32_t* __restrict__ block_of_row, uint64_t seed, uint32_t row_offset,
                      double* __restrict__ out_e, double* __restrict__ out_f, double* __restrict__ out_cos,
                      double* __restrict__ out_theta, double* __restrict__ out_sac, double* __restrict__ out_pos,
                      uint8_t* __restrict__ keep) {
    const uint64_t n = n_rows * (uint64_t)n_samples;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int nu = mode == 0 ? 2 : 3;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t r = i / n_samples;
        const uint32_t s = (uint32_t)(i - r * n_samples);
        double row[MR_NP];
#pragma unroll
        for (int k = 0; k < MR_NP; ++k) row[k] = __ldg(table + r * MR_NP + k);
        const bool live = row[MR_LIVE] != 0.0;
        double u_e = 0.5, u_c = 0.5, u_phi = 0.5;
        if (live) {
            if (uniforms) {
                const int32_t blk = __ldg(block_of_row + r);
                const double* ub = uniforms + ((size_t)blk * nu) * n_samples;
                u_e = __ldg(ub + s); u_c = __ldg(ub + n_samples + s);
                if (mode == 1) u_phi = __ldg(ub + 2 * (size_t)n_samples + s);
            } else {
                uniform_2(seed, kTagMeson3EC, (uint32_t)r + row_offset, s, u_e, u_c);
                if (mode == 1) u_phi = uniform_1(seed, kTagMeson3Phi, (uint32_t)r + row_offset, s);
            }
        }
        if (mode == 0) {
            double w;
            out_e[i] = meson3_sample_isotropic(u_e, u_c, row, K, w);
            out_f[i] = w;
            if (keep) keep[i] = 1;
        } else {
            Meson3DecayOut o;
            if (live) o = meson3_sample_decay(u_e, u_c, u_phi, row, K, energy_cut);
            else { o.e_lab = 0.0; o.cos_lab = 0.0; o.theta_z = 0.0; o.flux = 0.0; o.accept = 0; }
            out_e[i] = o.e_lab; out_f[i] = o.flux; out_cos[i] = o.cos_lab; out_theta[i] = o.theta_z;
            out_sac[i] = row[MR_SAC]; out_pos[i] = row[MR_POS];
            keep[i] = (live && (o.accept || !cut_on_solid_angle)) ? 1 : 0;        // :717-718
        }
    }
}

// dGammadEa on an array of energies (the classes' public method; also used by total_br on the host side)
__global__ void __launch_bounds__(kThreads)
meson3_dgamma_kernel(const double* __restrict__ ea, uint64_t n, double ml, Meson3Const K, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = meson3_dgamma_dea(__ldg(ea + i), ml, K);
}

__global__ void __launch_bounds__(kThreads)
meson3_m2_kernel(double m122, const double* __restrict__ m232, uint64_t n, double ml, Meson3Const K, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = m2_meson3body(m122, __ldg(m232 + i), ml, K, m2_threshold_factor(m122, K));
}

}  // namespace alpk

using namespace alpk;

ALPK_EXPORT int alpk_meson3_m2(double m122, const double* m232, uint64_t n, double lepton_mass, const alpk_meson3_t* h_par,
                               double* out, void* stream) {
    if (n == 0) return 0;
    if (!m232 || !h_par || !out) { set_error("alpk_meson3_m2: null argument"); return 2; }
    meson3_m2_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(m122, m232, n, lepton_mass, to_meson3(h_par), out);
    return check_launch("alpk_meson3_m2");
}

ALPK_EXPORT int alpk_meson3_rows(int mode, const double* mesons, uint64_t n_mesons, const uint32_t* row_meson,
                                 const double* row_ml, uint64_t n_rows, const alpk_meson3_t* h_par, double det_dist,
                                 double det_length, const double* u_pos, uint64_t seed, uint32_t meson_offset,
                                 double* row_table, void* stream) {
    if (mode != 0 && mode != 1) { set_error("alpk_meson3_rows: mode must be 0 (isotropic) or 1 (decay in flight)"); return 2; }
    if (n_rows == 0) return 0;
    if (!mesons || !row_meson || !row_ml || !h_par || !row_table || n_mesons == 0) { set_error("alpk_meson3_rows: null argument"); return 2; }
    meson3_rows_kernel<<<grid_for(n_rows, kThreads, 4), kThreads, 0, (cudaStream_t)stream>>>(
        mode, mesons, row_meson, row_ml, n_rows, to_meson3(h_par), det_dist, det_length, u_pos, seed, meson_offset, row_table);
    return check_launch("alpk_meson3_rows");
}

ALPK_EXPORT int alpk_meson3_samples(int mode, const double* row_table, uint64_t n_rows, uint32_t n_samples,
                                    const alpk_meson3_t* h_par, double energy_cut, int cut_on_solid_angle,
                                    const double* uniforms, const int32_t* block_of_row, uint64_t seed, uint32_t row_offset,
                                    double* out_energy, double* out_flux, double* out_cos, double* out_theta,
                                    double* out_solid_angle, double* out_decay_pos, uint8_t* keep, void* stream) {
    if (mode != 0 && mode != 1) { set_error("alpk_meson3_samples: mode must be 0 or 1"); return 2; }
    if (n_rows == 0 || n_samples == 0) return 0;
    if (!row_table || !h_par || !out_energy || !out_flux) { set_error("alpk_meson3_samples: null argument"); return 2; }
    if (mode == 1 && (!out_cos || !out_theta || !out_solid_angle || !out_decay_pos || !keep)) {
        set_error("alpk_meson3_samples: decay mode needs every output array"); return 2;
    }
    if (uniforms && !block_of_row) { set_error("alpk_meson3_samples: injected uniforms need block_of_row"); return 2; }
    const uint64_t n = n_rows * (uint64_t)n_samples;
    meson3_samples_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(
        mode, row_table, n_rows, n_samples, to_meson3(h_par), energy_cut, cut_on_solid_angle, uniforms, block_of_row, seed,
        row_offset, out_energy, out_flux, out_cos, out_theta, out_solid_angle, out_decay_pos, keep);
    return check_launch("alpk_meson3_samples");
}

ALPK_EXPORT int alpk_meson3_dgamma(const double* ea, uint64_t n, double lepton_mass, const alpk_meson3_t* h_par,
                                   double* out, void* stream) {
    if (n == 0) return 0;
    if (!ea || !h_par || !out) { set_error("alpk_meson3_dgamma: null argument"); return 2; }
    meson3_dgamma_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(ea, n, lepton_mass, to_meson3(h_par), out);
    return check_launch("alpk_meson3_dgamma");
}
