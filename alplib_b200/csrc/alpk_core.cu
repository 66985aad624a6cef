// libalpk core: error plumbing, propagate / event-weight kernels, table lookups, Primakoff and Compton production.
#include <stdarg.h>
#include <string.h>

#include "../../include/alpk.h"
#include "chain_kernel.cuh"

using namespace alp;

namespace alpk {

constexpr int kMaxTable = 512;

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;      // kernels launched through this library (bench.py reports it)

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
    __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED);
    return 0;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

// ---------------------------------------------------------------------------------------------------------------
// deterministic second stage of every reduction: one CTA sums the per-CTA partials in a fixed order
// ---------------------------------------------------------------------------------------------------------------
__global__ void finalize_sum_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double red[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
    v = block_sum(v, red);
    if (threadIdx.x == 0) *out = v;
}

int finalize_sum(const double* partials, int n, double* out, cudaStream_t st) {
    finalize_sum_kernel<<<1, 256, 0, st>>>(partials, n, out);
    return check_launch("finalize_sum");
}

// ---------------------------------------------------------------------------------------------------------------
// propagate  (fluxes.py:43-65, 159-162)
// ---------------------------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(kThreads)
propagate_kernel(const double* __restrict__ energy, const double* __restrict__ weight, uint64_t n, PropConst c,
                 float* __restrict__ d32, float* __restrict__ s32, double* __restrict__ d64o, double* __restrict__ s64o) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double d64, s64;
        if (FAST) propagate_sample_fast(__ldg(energy + i), __ldg(weight + i), c, d64, s64);
        else propagate_sample(__ldg(energy + i), __ldg(weight + i), c, d64, s64);
        d32[i] = FAST ? prop_cast_fast(d64, c) : prop_cast(d64, c);
        s32[i] = FAST ? prop_cast_fast(s64, c) : prop_cast(s64, c);
        if (d64o) d64o[i] = d64;
        if (s64o) s64o[i] = s64;
    }
}

// Fast mode, pair-vectorised: each thread takes two adjacent samples (16-byte loads, 8-byte float stores), the exp table
// sits in shared memory and the arithmetic of both samples is straight-line (see chain_pair); rare samples outside the
// branch-free domain are redone through propagate_sample_fast.  Needs 16-byte aligned arrays (the launcher checks).
template <bool SHORT>
__global__ void __launch_bounds__(kThreads)
propagate_pairs_kernel(const double* __restrict__ energy, const double* __restrict__ weight, uint64_t n, PropConst c,
                       float* __restrict__ d32, float* __restrict__ s32, double* __restrict__ d64o, double* __restrict__ s64o) {
    __shared__ double s_tab[alp::kExpTabSize];
    for (int i = threadIdx.x; i < alp::kExpTabSize; i += blockDim.x) s_tab[i] = kExpTabDev[i];
    __syncthreads();
    const uint64_t n_pairs = n >> 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += stride) {
        const double2 e = __ldg(reinterpret_cast<const double2*>(energy) + p);
        const double2 w = __ldg(reinterpret_cast<const double2*>(weight) + p);
        double d0, s0, d1, s1;
        bool ok = propagate_try_fast<SHORT>(e.x, w.x, c, d0, s0, s_tab);
        ok &= propagate_try_fast<SHORT>(e.y, w.y, c, d1, s1, s_tab);
        if (!ok) {
            propagate_sample_fast(e.x, w.x, c, d0, s0);
            propagate_sample_fast(e.y, w.y, c, d1, s1);
        }
        reinterpret_cast<float2*>(d32)[p] = make_float2(prop_cast_fast(d0, c), prop_cast_fast(d1, c));
        reinterpret_cast<float2*>(s32)[p] = make_float2(prop_cast_fast(s0, c), prop_cast_fast(s1, c));
        if (d64o) reinterpret_cast<double2*>(d64o)[p] = make_double2(d0, d1);
        if (s64o) reinterpret_cast<double2*>(s64o)[p] = make_double2(s0, s1);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {          // odd tail
        const uint64_t i = n - 1;
        double d64, s64;
        propagate_sample_fast(__ldg(energy + i), __ldg(weight + i), c, d64, s64);
        d32[i] = prop_cast_fast(d64, c);
        s32[i] = prop_cast_fast(s64, c);
        if (d64o) d64o[i] = d64;
        if (s64o) s64o[i] = s64;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// decays / scatter event weights + rate  (generators.py:41-46, 80-92)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
decays_kernel(const double* __restrict__ energy, const float* __restrict__ w32, uint64_t n, EventConst c,
              double* __restrict__ event_w, double* __restrict__ partials) {
    __shared__ double red[32];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ev = decay_event_weight(__ldg(energy + i), __ldg(w32 + i), c);
        if (event_w) event_w[i] = ev;
        acc += ev;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

struct ScatterConst {
    int kind;
    IComptonConst ic;
    double ma, ma2, pref, r02;
    double targets_per_area, mbm2;
};

// days*S_PER_DAY * (ntargets/det_area) * sigma(E) * METER_BY_MEV**2 * scatter_w32 * heaviside(E-thr, 1)
__global__ void __launch_bounds__(kThreads)
scatter_kernel(const double* __restrict__ energy, const float* __restrict__ w32, uint64_t n, ScatterConst sc,
               EventConst c, double* __restrict__ event_w, double* __restrict__ partials) {
    __shared__ double red[32];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const double lead = c.days_sec * sc.targets_per_area;       // Python floats: (days*S_PER_DAY)*(ntargets/area)
    double acc = 0.0;
    // four samples per trip with the loads issued first (one load in flight per thread left the kernel waiting on HBM
    // latency: 15 % of the bandwidth, long-scoreboard stalls dominating)
    constexpr int U = 4;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
        double ea[U];
        float w[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const uint64_t i = i0 + q * stride;
            ea[q] = i < n ? __ldg(energy + i) : 0.0;
            w[q] = i < n ? __ldg(w32 + i) : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const uint64_t i = i0 + q * stride;
            if (i >= n) break;
            double xs;
            if (sc.kind == 0) xs = icompton_sigma(ea[q], sc.ic);
            else if (sc.kind == 1) xs = iprimakoff_sigma(ea[q], sc.ma, sc.ma2, sc.pref, sc.r02);
            else if (sc.kind == 2) xs = icompton_sigma_fast(ea[q], sc.ic);
            else xs = iprimakoff_sigma_fast(ea[q], sc.ma, sc.ma2, sc.pref, sc.r02);
            const double ev = (((lead * xs) * sc.mbm2) * (double)w[q]) * heaviside(ea[q] - c.threshold, 1.0);
            if (event_w) event_w[i] = ev;
            acc += ev;
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// pair production a e- -> e- e+ e- (generators.py:33-39, photon_xs.py:95-125): tabulated cross section,
// heaviside(E-2me,0) * 10**interp(log10 E; left=-inf, right=fp[-1])
__global__ void __launch_bounds__(kThreads)
pair_kernel(const double* __restrict__ energy, const float* __restrict__ w32, uint64_t n,
            const double* __restrict__ log10_e, const double* __restrict__ log10_xs, int n_table,
            double lead, double mbm2, EventConst c, double* __restrict__ event_w, double* __restrict__ partials) {
    __shared__ double red[32];
    __shared__ double s_e[kMaxTable], s_xs[kMaxTable];
    for (int i = threadIdx.x; i < n_table; i += blockDim.x) { s_e[i] = __ldg(log10_e + i); s_xs[i] = __ldg(log10_xs + i); }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double ea = __ldg(energy + i);
        const double y = interp(log10(ea), s_e, s_xs, n_table, -INFINITY, s_xs[n_table - 1]);
        const double xs = heaviside(ea - 2 * kME, 0.0) * pow(10.0, y);            // photon_xs.py:122-125
        const double ev = (((lead * xs) * mbm2) * (double)__ldg(w32 + i)) * heaviside(ea - c.threshold, 1.0);
        if (event_w) event_w[i] = ev;
        acc += ev;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(kThreads)
sum_kernel(const double* __restrict__ x, uint64_t n, double* __restrict__ partials) {
    __shared__ double red[32];
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += __ldg(x + i);
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// table lookup + Primakoff rows + Compton rows (one thread per flux row; the XCOM table lives in shared memory)
// ---------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void stage_table(double* s_e, double* s_xs, const double* __restrict__ log10_e,
                                            const double* __restrict__ log10_xs, int n_table) {
    for (int i = threadIdx.x; i < n_table; i += blockDim.x) {
        s_e[i] = __ldg(log10_e + i);
        s_xs[i] = __ldg(log10_xs + i);
    }
    __syncthreads();
}

// Generic log-log table lookup of the tabulated cross-section classes (photon_xs.py:45-52, 117-128):
//   kind 0: 10**interp(log10 E; left = 0, right = 0)                                  AbsCrossSection.sigma_cm2 / sigma_mev
//   kind 1: heaviside(E - 2 me, 0) * 10**interp(log10 E; left = -inf, right = fp[-1])   (ALP)PairProdutionCrossSection
// The unit (cm2 or MeV^-2) is whatever the caller folded into log10_xs, exactly as the reference folds it.
__global__ void __launch_bounds__(kThreads)
table_sigma_kernel(int kind, const double* __restrict__ log10_e, const double* __restrict__ log10_xs, int n_table,
                   const double* __restrict__ energy, uint64_t n, double* __restrict__ out) {
    __shared__ double s_e[kMaxTable], s_xs[kMaxTable];
    stage_table(s_e, s_xs, log10_e, log10_xs, n_table);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = __ldg(energy + i);
        if (kind == 0) out[i] = pow(10.0, interp(log10(e), s_e, s_xs, n_table, 0.0, 0.0));
        else out[i] = heaviside(e - 2 * kME, 0.0) * pow(10.0, interp(log10(e), s_e, s_xs, n_table, -INFINITY, s_xs[n_table - 1]));
    }
}

__global__ void __launch_bounds__(kThreads)
abs_sigma_kernel(const double* __restrict__ log10_e, const double* __restrict__ log10_xs, int n_table,
                 const double* __restrict__ energy, uint64_t n, double* __restrict__ out) {
    __shared__ double s_e[kMaxTable], s_xs[kMaxTable];
    stage_table(s_e, s_xs, log10_e, log10_xs, n_table);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = abs_sigma_mev(__ldg(energy + i), s_e, s_xs, n_table);
}

__global__ void __launch_bounds__(kThreads)
primakoff_kernel(const double* __restrict__ eg, const double* __restrict__ phi, uint64_t n,
                 const double* __restrict__ log10_e, const double* __restrict__ log10_xs, int n_table,
                 double ma, double pref, double r02, double* __restrict__ out_e, double* __restrict__ out_f) {
    __shared__ double s_e[kMaxTable], s_xs[kMaxTable];
    stage_table(s_e, s_xs, log10_e, log10_xs, n_table);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = __ldg(eg + i);
        const double xs = primakoff_sigma(e, ma, pref, r02);               // fluxes.py:139
        const double br = xs / abs_sigma_mev(e, s_e, s_xs, n_table);       // :140
        out_e[i] = e;                                                      // :141
        out_f[i] = __ldg(phi + i) * br;                                    // :142
    }
}

// Primakoff production -> propagate -> decay event weights in one pass over the rows (one ALP sample per photon row): the
// arithmetic is the separate kernels' (primakoff_kernel, propagate_sample[_fast] + prop_cast, decay_event_weight), the grid
// and the accumulation order are decays_kernel's, so every array and the rate have the bits of the call-by-call schedule.
template <bool FAST>
__global__ void __launch_bounds__(kThreads)
primakoff_chain_kernel(const double* __restrict__ eg, const double* __restrict__ phi, uint64_t n,
                       const double* __restrict__ log10_e, const double* __restrict__ log10_xs, int n_table,
                       double ma, double pref, double r02, PropConst pc, EventConst ec,
                       double* __restrict__ out_e, double* __restrict__ out_f, float* __restrict__ d32o, float* __restrict__ s32o,
                       double* __restrict__ d64o, double* __restrict__ s64o, double* __restrict__ event_w,
                       double* __restrict__ partials) {
    __shared__ double s_e[kMaxTable], s_xs[kMaxTable];
    __shared__ double red[32];
    stage_table(s_e, s_xs, log10_e, log10_xs, n_table);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = __ldg(eg + i);
        const double xs = primakoff_sigma(e, ma, pref, r02);               // fluxes.py:139
        const double br = xs / abs_sigma_mev(e, s_e, s_xs, n_table);       // :140
        const double f = __ldg(phi + i) * br;                              // :142
        double d64, s64;
        if (FAST) propagate_sample_fast(e, f, pc, d64, s64); else propagate_sample(e, f, pc, d64, s64);
        const float d32 = FAST ? prop_cast_fast(d64, pc) : prop_cast(d64, pc);
        const float s32 = FAST ? prop_cast_fast(s64, pc) : prop_cast(s64, pc);
        const double ev = decay_event_weight(e, d32, ec);
        acc += ev;
        if (out_e) out_e[i] = e;
        if (out_f) out_f[i] = f;
        if (d32o) d32o[i] = d32;
        if (s32o) s32o[i] = s32;
        if (d64o) d64o[i] = d64;
        if (s64o) s64o[i] = s64;
        if (event_w) event_w[i] = ev;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(kThreads)
compton_rows_kernel(const double* __restrict__ eg, const double* __restrict__ phi, uint64_t n,
                    const double* __restrict__ log10_e, const double* __restrict__ log10_xs, int n_table,
                    ComptonGlobals g, double* __restrict__ table) {
    __shared__ double s_e[kMaxTable], s_xs[kMaxTable];
    stage_table(s_e, s_xs, log10_e, log10_xs, n_table);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double e = __ldg(eg + i);
        double row[CB_NP];
        compton_bin_params(e, __ldg(phi + i), abs_sigma_mev(e, s_e, s_xs, n_table), g, row);
#pragma unroll
        for (int k = 0; k < CB_NP; ++k) table[i * CB_NP + k] = row[k];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Compton production channel for the generic chain kernel (chain_kernel.cuh)
// ---------------------------------------------------------------------------------------------------------------
struct ComptonProd {
    static constexpr int NP = CB_NP;
    static constexpr uint32_t kTag = kTagComptonEa;
    using Globals = ComptonGlobals;
    __device__ static __forceinline__ uint64_t uidx(uint64_t row, uint32_t s, uint32_t n) { return row * n + s; }
    __device__ static __forceinline__ double sample(double u, const double* b, const Globals& g, double& flux) {
        return compton_sample(u, b, g, flux);
    }
    __device__ static __forceinline__ double sample_fast(double u, const double* b, const Globals& g, double& flux, const double* exp_tab) {
        return compton_sample_fast(u, b, g, flux);
    }
};

PropConst to_prop(const alpk_prop_t* p) {
    PropConst c;
    c.ma = p->ma; c.ma2 = p->ma2; c.width = p->width; c.rescale = p->rescale;
    c.neg_dist_mbm = p->neg_dist_mbm; c.neg_len_mbm = p->neg_len_mbm; c.geom = p->geom; c.geom32 = p->geom32;
    c.apply_geom = p->apply_geom; c.geom_f64 = p->geom_f64; c.use_tof = p->use_tof; c.tof_light = p->tof_light;
    c.det_dist = p->det_dist; c.timing_window = p->timing_window;
    c.fast = p->fast; c.k1 = p->k1; c.k2 = p->k2;
    prop_derive(c);
    return c;
}

EventConst to_event(const alpk_event_t* e) {
    EventConst c;
    c.days_sec = e->days_sec; c.days_sec32 = e->days_sec32; c.days_f64 = e->days_f64; c.threshold = e->threshold;
    return c;
}

static ComptonGlobals to_compton(const alpk_compton_t* p) {
    ComptonGlobals g;
    g.ma = p->ma; g.ma2 = p->ma2; g.aa = p->aa; g.z = p->z; g.n_samples = p->n_samples;
    return g;
}

}  // namespace alpk

using namespace alpk;

// ===================================================================================================================
// C ABI
// ===================================================================================================================
ALPK_EXPORT int alpk_version(void) { return ALPK_VERSION; }

ALPK_EXPORT const char* alpk_last_error(void) { return g_err; }

ALPK_EXPORT uint64_t alpk_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

ALPK_EXPORT int alpk_device_info(int device, char* h_name, int name_len, int* h_sm_count, int* h_cc_major, int* h_cc_minor) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e)); return 1; }
    if (h_name && name_len > 0) { strncpy(h_name, p.name, (size_t)name_len - 1); h_name[name_len - 1] = 0; }
    if (h_sm_count) *h_sm_count = p.multiProcessorCount;
    if (h_cc_major) *h_cc_major = p.major;
    if (h_cc_minor) *h_cc_minor = p.minor;
    return 0;
}

// ---- minimal device-memory helpers so that a caller without PyTorch / the CUDA runtime can drive the library -----------
ALPK_EXPORT int alpk_set_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e)); return 1; }
    return 0;
}

ALPK_EXPORT int alpk_malloc(void** d_ptr, uint64_t bytes) {
    if (!d_ptr) { set_error("alpk_malloc: null argument"); return 2; }
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) { set_error("cudaMalloc(%llu): %s", (unsigned long long)bytes, cudaGetErrorString(e)); return 1; }
    return 0;
}

ALPK_EXPORT int alpk_free(void* d_ptr) {
    cudaError_t e = cudaFree(d_ptr);
    if (e != cudaSuccess) { set_error("cudaFree: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

ALPK_EXPORT int alpk_memcpy_h2d(void* d_dst, const void* h_src, uint64_t bytes, void* stream) {
    cudaError_t e = cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("alpk_memcpy_h2d: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

ALPK_EXPORT int alpk_memcpy_d2h(void* h_dst, const void* d_src, uint64_t bytes, void* stream) {
    cudaError_t e = cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("alpk_memcpy_d2h: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

// Host-side staging of a flux table for the production kernels (fluxes.py:136 / :214-216 row filter + SoA split):
// one pass over the caller's row-major [n_rows][row_stride] table gathers the kept rows' energies, weights and global
// row ids into the pinned staging buffer, then ONE asynchronous H2D copy moves the block.
ALPK_EXPORT int alpk_stage_flux_rows(const double* h_table, uint64_t n_rows, uint32_t row_stride, int rule, double cut,
                                     uint32_t row_offset, void* h_pinned, uint64_t pinned_bytes, void* d_dst,
                                     uint64_t* h_n_kept, uint64_t* h_off_weight, uint64_t* h_off_ids, void* stream) {
    if (!h_n_kept || !h_off_weight || !h_off_ids) { set_error("alpk_stage_flux_rows: null output"); return 2; }
    *h_n_kept = 0; *h_off_weight = 0; *h_off_ids = 0;
    if (n_rows == 0) return 0;
    if (!h_table || !h_pinned || !d_dst || row_stride < 2) { set_error("alpk_stage_flux_rows: null argument"); return 2; }
    if (rule != 0 && rule != 1) { set_error("alpk_stage_flux_rows: rule must be 0 (E >= m_a) or 1 (Compton threshold)"); return 2; }
    const uint64_t cap = (n_rows * 8 + 255) & ~255ull;                    // per-array capacity in the staging block
    if (pinned_bytes < 2 * cap + n_rows * 4) { set_error("alpk_stage_flux_rows: staging buffer too small"); return 2; }
    double* e = (double*)h_pinned;
    double* w = (double*)((char*)h_pinned + cap);
    uint32_t* ids = (uint32_t*)((char*)h_pinned + 2 * cap);
    const double me = alp::kME, me2 = alp::kME * alp::kME;
    // branch-free compaction: every row is written at the cursor, the cursor advances only for kept rows (the staging
    // block has room for all n_rows)
    uint64_t k = 0;
    if (rule == 0) {
        for (uint64_t r = 0; r < n_rows; ++r) {
            const double en = h_table[r * row_stride];
            e[k] = en; w[k] = h_table[r * row_stride + 1]; ids[k] = (uint32_t)r + row_offset;
            k += !(en < cut);                                             // fluxes.py:136   if photon[0] < self.ma: return
        }
    } else {
        for (uint64_t r = 0; r < n_rows; ++r) {
            const double en = h_table[r * row_stride];
            e[k] = en; w[k] = h_table[r * row_stride + 1]; ids[k] = (uint32_t)r + row_offset;
            k += !(((2 * me) * en + me2) < cut);                          // :214-215  s < (M_E + ma)**2
        }
    }
    *h_n_kept = k;
    if (k == 0) return 0;
    // compact layout on both sides: [E: k][pad][w: k][pad][ids: k], arrays 256-byte aligned
    const uint64_t seg = (k * 8 + 255) & ~255ull;
    if (seg != cap) {
        memmove((char*)h_pinned + seg, w, k * 8);
        memmove((char*)h_pinned + 2 * seg, ids, k * 4);
    }
    *h_off_weight = seg; *h_off_ids = 2 * seg;
    cudaError_t err = cudaMemcpyAsync(d_dst, h_pinned, 2 * seg + k * 4, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (err != cudaSuccess) { set_error("alpk_stage_flux_rows: %s", cudaGetErrorString(err)); return 1; }
    return 0;
}

ALPK_EXPORT int alpk_sync(void* stream) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("alpk_sync: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

ALPK_EXPORT int alpk_propagate(const double* energy, const double* weight, uint64_t n, const alpk_prop_t* h_prop,
                               float* decay_w32, float* scatter_w32, double* d64, double* s64, void* stream) {
    if (!h_prop || !decay_w32 || !scatter_w32) { set_error("alpk_propagate: null argument"); return 2; }
    if (n == 0) return 0;
    if (!energy || !weight) { set_error("alpk_propagate: null input"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    const uintptr_t align = (uintptr_t)energy | (uintptr_t)weight | (uintptr_t)decay_w32 | (uintptr_t)scatter_w32 |
                            (uintptr_t)d64 | (uintptr_t)s64;
    if (h_prop->fast && (align & 15) == 0) {
        const PropConst c = to_prop(h_prop);
        if (c.short_lived) propagate_pairs_kernel<true><<<grid_for(n, kThreads * 2 * 4, 8), kThreads, 0, st>>>(energy, weight, n, c, decay_w32, scatter_w32, d64, s64);
        else propagate_pairs_kernel<false><<<grid_for(n, kThreads * 2 * 4, 8), kThreads, 0, st>>>(energy, weight, n, c, decay_w32, scatter_w32, d64, s64);
    } else if (h_prop->fast)
        propagate_kernel<true><<<grid_for(n, kThreads, 8), kThreads, 0, st>>>(energy, weight, n, to_prop(h_prop), decay_w32,
                                                                             scatter_w32, d64, s64);
    else
        propagate_kernel<false><<<grid_for(n, kThreads, 8), kThreads, 0, st>>>(energy, weight, n, to_prop(h_prop), decay_w32,
                                                                              scatter_w32, d64, s64);
    return check_launch("alpk_propagate");
}

ALPK_EXPORT int alpk_decays(const double* energy, const float* decay_w32, uint64_t n, const alpk_event_t* h_ev,
                            double* event_w, double* rate, double* scratch, void* stream) {
    if (!h_ev || !rate || !scratch) { set_error("alpk_decays: null argument"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { cudaMemsetAsync(rate, 0, sizeof(double), st); return check_launch("alpk_decays(memset)"); }
    if (!energy || !decay_w32) { set_error("alpk_decays: null input"); return 2; }
    const int grid = grid_for(n, kThreads * 4, 8);
    decays_kernel<<<grid, kThreads, 0, st>>>(energy, decay_w32, n, to_event(h_ev), event_w, scratch);
    if (check_launch("alpk_decays")) return 1;
    return finalize_sum(scratch, grid, rate, st);
}

ALPK_EXPORT int alpk_scatter_events(int kind, const double* energy, const float* scatter_w32, uint64_t n,
                                    const double* h_xs, double prefactor, double mbm2, const alpk_event_t* h_ev,
                                    double* event_w, double* rate, double* scratch, void* stream) {
    if (!h_ev || !rate || !scratch || !h_xs) { set_error("alpk_scatter_events: null argument"); return 2; }
    if (kind < 0 || kind > 3) { set_error("alpk_scatter_events: unknown kind %d", kind); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { cudaMemsetAsync(rate, 0, sizeof(double), st); return check_launch("alpk_scatter_events(memset)"); }
    if (!energy || !scatter_w32) { set_error("alpk_scatter_events: null input"); return 2; }
    ScatterConst sc;
    memset(&sc, 0, sizeof(sc));
    sc.kind = kind;
    if (kind == 0 || kind == 2) {
        sc.ic.ma = h_xs[0]; sc.ic.ma2 = h_xs[1]; sc.ic.ma4 = h_xs[2]; sc.ic.ethr = h_xs[3]; sc.ic.zfac = h_xs[4]; sc.ic.c2 = h_xs[5];
    } else {
        sc.ma = h_xs[0]; sc.ma2 = h_xs[1]; sc.pref = h_xs[2]; sc.r02 = h_xs[3];
    }
    sc.targets_per_area = prefactor; sc.mbm2 = mbm2;
    const int grid = grid_for(n, kThreads * 2, 8);
    scatter_kernel<<<grid, kThreads, 0, st>>>(energy, scatter_w32, n, sc, to_event(h_ev), event_w, scratch);
    if (check_launch("alpk_scatter_events")) return 1;
    return finalize_sum(scratch, grid, rate, st);
}

ALPK_EXPORT int alpk_pair_events(const double* energy, const float* scatter_w32, uint64_t n, const double* log10_e,
                                 const double* log10_xs, int n_table, double lead, double mbm2, const alpk_event_t* h_ev,
                                 double* event_w, double* rate, double* scratch, void* stream) {
    if (!h_ev || !rate || !scratch || !log10_e || !log10_xs) { set_error("alpk_pair_events: null argument"); return 2; }
    if (n_table < 1 || n_table > kMaxTable) { set_error("alpk_pair_events: table size %d not in [1,%d]", n_table, kMaxTable); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { cudaMemsetAsync(rate, 0, sizeof(double), st); return 0; }
    if (!energy || !scatter_w32) { set_error("alpk_pair_events: null input"); return 2; }
    const int grid = grid_for(n, kThreads * 2, 8);
    pair_kernel<<<grid, kThreads, 0, st>>>(energy, scatter_w32, n, log10_e, log10_xs, n_table, lead, mbm2, to_event(h_ev), event_w, scratch);
    if (check_launch("alpk_pair_events")) return 1;
    return finalize_sum(scratch, grid, rate, st);
}

ALPK_EXPORT int alpk_sum(const double* x, uint64_t n, double* out, double* scratch, void* stream) {
    if (!out || !scratch) { set_error("alpk_sum: null argument"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) { cudaMemsetAsync(out, 0, sizeof(double), st); return check_launch("alpk_sum(memset)"); }
    if (!x) { set_error("alpk_sum: null input"); return 2; }
    const int grid = grid_for(n, kThreads * 8, 8);
    sum_kernel<<<grid, kThreads, 0, st>>>(x, n, scratch);
    if (check_launch("alpk_sum")) return 1;
    return finalize_sum(scratch, grid, out, st);
}

ALPK_EXPORT int alpk_table_sigma(int kind, const double* log10_e, const double* log10_xs, int n_table, const double* energy,
                                 uint64_t n, double* sigma, void* stream) {
    if (kind != 0 && kind != 1) { set_error("alpk_table_sigma: kind must be 0 or 1"); return 2; }
    if (n_table < 1 || n_table > kMaxTable) { set_error("alpk_table_sigma: table size %d not in [1,%d]", n_table, kMaxTable); return 2; }
    if (n == 0) return 0;
    if (!log10_e || !log10_xs || !energy || !sigma) { set_error("alpk_table_sigma: null argument"); return 2; }
    table_sigma_kernel<<<grid_for(n, kThreads, 4), kThreads, 0, (cudaStream_t)stream>>>(kind, log10_e, log10_xs, n_table, energy, n, sigma);
    return check_launch("alpk_table_sigma");
}

ALPK_EXPORT int alpk_abs_sigma_mev(const double* log10_e, const double* log10_xs, int n_table, const double* energy,
                                   uint64_t n, double* sigma_mev, void* stream) {
    if (n_table < 1 || n_table > kMaxTable) { set_error("alpk_abs_sigma_mev: table size %d not in [1,%d]", n_table, kMaxTable); return 2; }
    if (n == 0) return 0;
    if (!log10_e || !log10_xs || !energy || !sigma_mev) { set_error("alpk_abs_sigma_mev: null argument"); return 2; }
    abs_sigma_kernel<<<grid_for(n, kThreads, 4), kThreads, 0, (cudaStream_t)stream>>>(log10_e, log10_xs, n_table, energy, n, sigma_mev);
    return check_launch("alpk_abs_sigma_mev");
}

ALPK_EXPORT int alpk_primakoff(const double* e_gamma, const double* phi_gamma, uint64_t n_rows,
                               const double* log10_e, const double* log10_xs, int n_table,
                               double ma, double pref, double r02, double* out_energy, double* out_flux, void* stream) {
    if (n_table < 1 || n_table > kMaxTable) { set_error("alpk_primakoff: table size %d not in [1,%d]", n_table, kMaxTable); return 2; }
    if (n_rows == 0) return 0;
    if (!e_gamma || !phi_gamma || !log10_e || !log10_xs || !out_energy || !out_flux) { set_error("alpk_primakoff: null argument"); return 2; }
    primakoff_kernel<<<grid_for(n_rows, kThreads, 4), kThreads, 0, (cudaStream_t)stream>>>(
        e_gamma, phi_gamma, n_rows, log10_e, log10_xs, n_table, ma, pref, r02, out_energy, out_flux);
    return check_launch("alpk_primakoff");
}

ALPK_EXPORT int alpk_primakoff_chain(const double* e_gamma, const double* phi_gamma, uint64_t n_rows,
                                     const double* log10_e, const double* log10_xs, int n_table,
                                     double ma, double pref, double r02, const alpk_prop_t* h_prop, const alpk_event_t* h_ev,
                                     double* out_energy, double* out_flux, float* decay_w32, float* scatter_w32,
                                     double* decay_w64, double* scatter_w64, double* event_w, double* rate, double* scratch,
                                     void* stream) {
    if (n_table < 1 || n_table > kMaxTable) { set_error("alpk_primakoff_chain: table size %d not in [1,%d]", n_table, kMaxTable); return 2; }
    if (!h_prop || !h_ev || !rate || !scratch) { set_error("alpk_primakoff_chain: null argument"); return 2; }
    cudaStream_t st = (cudaStream_t)stream;
    if (n_rows == 0) { cudaMemsetAsync(rate, 0, sizeof(double), st); return check_launch("alpk_primakoff_chain(memset)"); }
    if (!e_gamma || !phi_gamma || !log10_e || !log10_xs) { set_error("alpk_primakoff_chain: null input"); return 2; }
    const int grid = grid_for(n_rows, kThreads * 4, 8);                    // decays_kernel's grid: the same rate bits
    if (h_prop->fast)
        primakoff_chain_kernel<true><<<grid, kThreads, 0, st>>>(e_gamma, phi_gamma, n_rows, log10_e, log10_xs, n_table, ma, pref, r02,
                                                               to_prop(h_prop), to_event(h_ev), out_energy, out_flux, decay_w32,
                                                               scatter_w32, decay_w64, scatter_w64, event_w, scratch);
    else
        primakoff_chain_kernel<false><<<grid, kThreads, 0, st>>>(e_gamma, phi_gamma, n_rows, log10_e, log10_xs, n_table, ma, pref, r02,
                                                                to_prop(h_prop), to_event(h_ev), out_energy, out_flux, decay_w32,
                                                                scatter_w32, decay_w64, scatter_w64, event_w, scratch);
    if (check_launch("alpk_primakoff_chain")) return 1;
    return finalize_sum(scratch, grid, rate, st);
}

ALPK_EXPORT int alpk_compton_rows(const double* e_gamma, const double* phi_gamma, uint64_t n_rows,
                                  const double* log10_e, const double* log10_xs, int n_table,
                                  const alpk_compton_t* h_par, double* row_table, void* stream) {
    static_assert(CB_NP == ALPK_COMPTON_NP, "header/table layout mismatch");
    if (n_table < 1 || n_table > kMaxTable) { set_error("alpk_compton_rows: table size %d not in [1,%d]", n_table, kMaxTable); return 2; }
    if (!h_par) { set_error("alpk_compton_rows: null argument"); return 2; }
    if (n_rows == 0) return 0;
    if (!e_gamma || !phi_gamma || !log10_e || !log10_xs || !row_table) { set_error("alpk_compton_rows: null argument"); return 2; }
    compton_rows_kernel<<<grid_for(n_rows, kThreads, 4), kThreads, 0, (cudaStream_t)stream>>>(
        e_gamma, phi_gamma, n_rows, log10_e, log10_xs, n_table, to_compton(h_par), row_table);
    return check_launch("alpk_compton_rows");
}

ALPK_EXPORT int alpk_compton_samples(const double* row_table, uint64_t n_rows, uint32_t n_samples, const uint32_t* row_ids,
                                     const alpk_compton_t* h_par, const double* uniforms, uint64_t seed,
                                     double* out_energy, double* out_flux,
                                     const alpk_prop_t* h_prop, float* decay_w32, float* scatter_w32,
                                     double* decay_w64, double* scatter_w64,
                                     const alpk_event_t* h_ev, double* event_w, double* rate, double* scratch,
                                     const double* hist_edges, int n_edges, double* hist, void* stream) {
    if (!h_par) { set_error("alpk_compton_samples: null argument"); return 2; }
    return launch_chain<ComptonProd>("alpk_compton_samples", row_table, n_rows, n_samples, row_ids, to_compton(h_par),
                                     h_par->fast, h_par->rng_rounds, uniforms, seed, out_energy, out_flux, h_prop, decay_w32, scatter_w32, decay_w64,
                                     scatter_w64, h_ev, event_w, rate, scratch, hist_edges, n_edges, hist, stream);
}

// ---------------------------------------------------------------------------------------------------------------
// FP64 issue-rate microbenchmark (roofline denominator measured on the box; MEASURED_PEAKS.json has no FP64 figure)
// ---------------------------------------------------------------------------------------------------------------
namespace alpk {
__global__ void __launch_bounds__(256)
fp64_peak_kernel(uint64_t iters, double a, double b, double* __restrict__ out) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (uint64_t i = 0; i < iters; ++i) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // never true; keeps the chains alive
}
// FP64 exp throughput (SURVEY section 8d asks for it next to the DFMA peak): four independent evaluations per trip of
// either the library exp (kind 0) or the fast-mode table exp with the table in shared memory (kind 1), arguments in [-20, 0]
template <int KIND>
__global__ void __launch_bounds__(256)
exp_peak_kernel(uint64_t iters, double* __restrict__ out) {
    __shared__ double s_tab[alp::kExpTabSize];
    for (int i = threadIdx.x; i < alp::kExpTabSize; i += blockDim.x) s_tab[i] = alp::kExpTabDev[i];
    __syncthreads();
    double x0 = -1e-3 * threadIdx.x - 0.5, x1 = x0 - 1.0, x2 = x0 - 2.0, x3 = x0 - 3.0, acc = 0.0;
    for (uint64_t i = 0; i < iters; ++i) {
        const double e0 = KIND ? alp::exp_fast_core(x0, s_tab) : exp(x0), e1 = KIND ? alp::exp_fast_core(x1, s_tab) : exp(x1);
        const double e2 = KIND ? alp::exp_fast_core(x2, s_tab) : exp(x2), e3 = KIND ? alp::exp_fast_core(x3, s_tab) : exp(x3);
        acc += (e0 + e1) + (e2 + e3);
        x0 = x0 * 0.999 - e0 * 1e-3; x1 = x1 * 0.999 - e1 * 1e-3; x2 = x2 * 0.999 - e2 * 1e-3; x3 = x3 * 0.999 - e3 * 1e-3;
        x0 = x0 < -20.0 ? x0 + 19.0 : x0; x1 = x1 < -20.0 ? x1 + 19.0 : x1; x2 = x2 < -20.0 ? x2 + 19.0 : x2; x3 = x3 < -20.0 ? x3 + 19.0 : x3;
    }
    if (acc == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
}  // namespace alpk

// launches `blocks` CTAs of 256 threads, each thread evaluating 4*iters exponentials (kind 0: library exp, 1: exp_fast_core)
ALPK_EXPORT int alpk_exp_peak(int kind, uint64_t iters, int blocks, double* out, void* stream) {
    if (!out || blocks < 1 || (kind != 0 && kind != 1)) { set_error("alpk_exp_peak: bad argument"); return 2; }
    if (kind) exp_peak_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out);
    else exp_peak_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, out);
    return check_launch("alpk_exp_peak");
}

// launches `blocks` CTAs of 256 threads, each thread doing 8*iters dependent-chain DFMAs (16*iters flops)
ALPK_EXPORT int alpk_fp64_peak(uint64_t iters, int blocks, double* out, void* stream) {
    if (!out || blocks < 1) { set_error("alpk_fp64_peak: bad argument"); return 2; }
    fp64_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 0.999999, 1e-7, out);
    return check_launch("alpk_fp64_peak");
}
