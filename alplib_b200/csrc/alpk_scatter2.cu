// libalpk: generic 2 -> 2 scattering Monte Carlo (Scatter2to2MC, cross_section_mc.py:148-275) with the matrix elements of
// matrix_element.py:629-734, and direct evaluation of those matrix elements (the reference classes are callable).
#include <string.h>

#include "../../include/alpk.h"
#include "alp_math4.cuh"
#include "common.cuh"

using namespace alp;

namespace alpk {

constexpr uint32_t kTagScatter2 = 8;      // Philox stream tag (4th counter word) of the 2 -> 2 scatter draws

static int to_m2(const alpk_m2_t* h, M2Const& c) {
    if (h->model < 0 || h->model > 4) { set_error("unknown matrix-element model %d", h->model); return 2; }
    c.model = h->model; c.ma2 = h->ma2; c.ma4 = h->ma4; c.mN2 = h->mN2; c.mN4 = h->mN4; c.z = h->z; c.ff_a2 = h->ff_a2;
    c.pref = h->pref; c.me2 = h->me2; c.me4 = h->me4; c.tmax = h->tmax;
    return 0;
}

__global__ void __launch_bounds__(kThreads)
m2_eval_kernel(M2Const c, const double* __restrict__ s, uint64_t s_stride, const double* __restrict__ t, uint64_t n,
               double den, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = m2_eval(c, __ldg(s + i * s_stride), __ldg(t + i));
        out[i] = den != 0.0 ? v / den : v;
    }
}

__global__ void __launch_bounds__(kThreads)
atomic_ff2_kernel(const double* __restrict__ q, uint64_t n, double z, double a2, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = atomic_elastic_ff2(__ldg(q + i), z, a2);
}

// One thread per sample.  Outputs are component planes [4][n] / [3][n] (coalesced); each may be NULL.
template <bool INJECT>
__global__ void __launch_bounds__(kThreads)
scatter2_kernel(Scatter2Const c, uint64_t n, const double* __restrict__ uniforms, uint64_t seed, uint32_t stream_id,
                double* __restrict__ t_out, double* __restrict__ w_out, double* __restrict__ p3cm, double* __restrict__ p3lab,
                double* __restrict__ p4lab, double* __restrict__ cosine, double* __restrict__ v3cm, double* __restrict__ v3lab,
                double* __restrict__ v4lab) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double u1, u2;
        if (INJECT) { u1 = __ldg(uniforms + i); u2 = __ldg(uniforms + n + i); }      // phi block, then theta block
        else uniform_2(seed, kTagScatter2, stream_id, i, u1, u2);
        Scatter2Out o;
        scatter2_sample(u1, u2, c, o);
        if (t_out) t_out[i] = o.t;
        if (w_out) w_out[i] = o.w;
        if (cosine) cosine[i] = o.cosine;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (p3cm) p3cm[(uint64_t)k * n + i] = o.p3cm[k];
            if (p3lab) p3lab[(uint64_t)k * n + i] = o.p3lab[k];
            if (p4lab) p4lab[(uint64_t)k * n + i] = o.p4lab[k];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {                                               // get_3velocity: p_k / p_0
            if (v3cm) v3cm[(uint64_t)k * n + i] = o.p3cm[k + 1] / o.p3cm[0];
            if (v3lab) v3lab[(uint64_t)k * n + i] = o.p3lab[k + 1] / o.p3lab[0];
            if (v4lab) v4lab[(uint64_t)k * n + i] = o.p4lab[k + 1] / o.p4lab[0];
        }
    }
}

}  // namespace alpk

using namespace alpk;

ALPK_EXPORT int alpk_m2_eval(const alpk_m2_t* h_m2, const double* s, uint64_t s_stride, const double* t, uint64_t n,
                             double den, double* out, void* stream) {
    if (!h_m2 || !out) { set_error("alpk_m2_eval: null argument"); return 2; }
    if (n == 0) return 0;
    if (!s || !t) { set_error("alpk_m2_eval: null input"); return 2; }
    M2Const c;
    if (to_m2(h_m2, c)) return 2;
    m2_eval_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(c, s, s_stride, t, n, den, out);
    return check_launch("alpk_m2_eval");
}

ALPK_EXPORT int alpk_atomic_ff2(const double* q, uint64_t n, double z, double a2, double* out, void* stream) {
    if (n == 0) return 0;
    if (!q || !out) { set_error("alpk_atomic_ff2: null argument"); return 2; }
    atomic_ff2_kernel<<<grid_for(n, kThreads, 8), kThreads, 0, (cudaStream_t)stream>>>(q, n, z, a2, out);
    return check_launch("alpk_atomic_ff2");
}

ALPK_EXPORT int alpk_scatter2to2(const alpk_scatter2_t* h_par, uint64_t n_samples, const double* uniforms, uint64_t seed,
                                 uint32_t stream_id, double* t_out, double* weights, double* p3_cm, double* p3_lab,
                                 double* p4_lab, double* cosine_lab, double* v3_cm, double* v3_lab, double* v4_lab,
                                 void* stream) {
    if (!h_par) { set_error("alpk_scatter2to2: null argument"); return 2; }
    if (n_samples == 0) return 0;
    Scatter2Const c;
    memset(&c, 0, sizeof(c));
    if (to_m2(&h_par->m2, c.m2)) return 2;
    c.s = h_par->s; c.msum = h_par->msum; c.pp = h_par->pp; c.ee = h_par->ee; c.p3_cm = h_par->p3_cm; c.e3_cm = h_par->e3_cm;
    c.vol = h_par->vol; c.p1_cm = h_par->p1_cm; c.n_samples = h_par->n_samples; c.den = h_par->den;
    for (int k = 0; k < 16; ++k) c.mat[k] = h_par->mat[k];
    for (int k = 0; k < 4; ++k) c.p12[k] = h_par->p12[k];
    c.log_sampling = h_par->log_sampling;
    const int grid = grid_for(n_samples, kThreads, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (uniforms) scatter2_kernel<true><<<grid, kThreads, 0, st>>>(c, n_samples, uniforms, seed, stream_id, t_out, weights, p3_cm,
                                                                   p3_lab, p4_lab, cosine_lab, v3_cm, v3_lab, v4_lab);
    else scatter2_kernel<false><<<grid, kThreads, 0, st>>>(c, n_samples, uniforms, seed, stream_id, t_out, weights, p3_cm, p3_lab,
                                                           p4_lab, cosine_lab, v3_cm, v3_lab, v4_lab);
    return check_launch("alpk_scatter2to2");
}
