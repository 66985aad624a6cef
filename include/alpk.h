/*
 * alpk.h -- C ABI of libalpk.so, the sm_100a CUDA implementation of alplib's Monte-Carlo flux-and-event path.
 *
 * The reference (athompson-git/alplib) is pure Python and has no FFI; its boundary for this path is the Python
 * class API (fluxes.py / generators.py).  `alplib_b200` keeps that API and binds the entry points below with
 * ctypes (INTEGRATION.md shows the stub a reference maintainer would add).  Each entry point cites the reference
 * code it replaces (file:line relative to the reference repository root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (cudaMalloc'd / torch CUDA tensor) unless its name starts with `h_`;
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it; nothing synchronises;
 *   - return value 0 = OK, non-zero = error, text from alpk_last_error() (thread local);
 *   - nothing is allocated or freed across the boundary; `scratch` arguments are caller-owned device buffers of
 *     at least ALPK_SCRATCH_DOUBLES doubles (8 MiB) used for the deterministic two-stage reductions (block / unit partial
 *     sums, then a fixed-order finalize) -- one scratch per stream: launches that share a scratch must be stream-ordered;
 *   - per-sample output arrays must be 16-byte aligned (8-byte for the float32 arrays): the kernels store pairs of samples
 *     as vectors; a misaligned pointer is rejected with rc 2 (alpk_propagate alone falls back to scalar stores);
 *   - all arithmetic is IEEE FP64 in the reference's operation order (library built with -fmad=false), weights
 *     are rounded to float32 exactly where the reference rounds them;
 *   - `uniforms` arguments: NULL = free-running Philox4x32-10 keyed by (seed, stream tag, row id, sample);
 *     non-NULL = injected U[0,1) variates in the reference's draw order (parity mode).
 */
#ifndef ALPK_H
#define ALPK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALPK_VERSION 1
#define ALPK_SCRATCH_DOUBLES 1048576

/* number of doubles per row of the per-row parameter tables */
#define ALPK_COMPTON_NP 12
#define ALPK_PARENT_NP 6
#define ALPK_PARENT_GENERAL_NP 20
#define ALPK_BREM_NP 6
#define ALPK_PAIRANN_NP 11
#define ALPK_GL_POINTS 32

int alpk_version(void);
const char* alpk_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t alpk_launch_count(void);
/* writes name, SM count, compute capability of `device`; used by the loader to refuse non-sm_100 devices */
int alpk_device_info(int device, char* h_name, int name_len, int* h_sm_count, int* h_cc_major, int* h_cc_minor);

/* ---- device-memory helpers: enough for a caller without PyTorch or the CUDA runtime API to drive the library
 * (the Python package itself uses torch tensors for device memory and passes their pointers). -------------------- */
int alpk_set_device(int device);
int alpk_malloc(void** d_ptr, uint64_t bytes);
int alpk_free(void* d_ptr);
int alpk_memcpy_h2d(void* d_dst, const void* h_src, uint64_t bytes, void* stream);
int alpk_memcpy_d2h(void* h_dst, const void* d_src, uint64_t bytes, void* stream);
int alpk_sync(void* stream);

/* ---- host-side staging of a flux table (fluxes.py:136 and :214-216: the per-row `return` of simulate_single) -------
 * h_table: row-major [n_rows][row_stride] doubles (energy, weight, ...).  rule 0 keeps rows with !(E < cut) (Primakoff,
 * cut = m_a); rule 1 keeps rows with !(2 M_E E + M_E^2 < cut) (Compton, cut = (M_E + m_a)^2).  The kept rows' energies,
 * weights and global row ids (row + row_offset, uint32) are gathered into the pinned block h_pinned (>= 2*align256(8 n_rows)
 * + 4 n_rows bytes) and copied with one cudaMemcpyAsync to d_dst (same capacity): energies at byte 0, weights at
 * *h_off_weight, ids at *h_off_ids.  The caller must not rewrite h_pinned before the copy has completed. */
int alpk_stage_flux_rows(const double* h_table, uint64_t n_rows, uint32_t row_stride, int rule, double cut,
                         uint32_t row_offset, void* h_pinned, uint64_t pinned_bytes, void* d_dst,
                         uint64_t* h_n_kept, uint64_t* h_off_weight, uint64_t* h_off_ids, void* stream);

/* ---- propagate: fluxes.py:43-65 (AxionFlux.propagate) + fluxes.py:159-162 (isotropic acceptance) ------------ */
typedef struct {
    double ma, ma2;          /* axion mass, ma**2 */
    double width;            /* decay width [MeV]; <= 0 -> infinite lifetime (fluxes.py:51) */
    double rescale;          /* rescale_factor, (g'/g)^2 for propagate(new_coupling) */
    double neg_dist_mbm;     /* -det_dist / METER_BY_MEV */
    double neg_len_mbm;      /* -det_length / METER_BY_MEV */
    double geom;             /* det_area / (4 pi det_dist^2) */
    float  geom32;           /* float32(geom) */
    int    apply_geom;       /* is_isotropic */
    int    geom_f64;         /* acceptance factor is an np.float64 (multiply in f64, then cast) */
    int    use_tof;          /* time_of_flight_cut is not None (fluxes.py:54-58) */
    double tof_light;        /* det_dist / (1e-2*C_LIGHT) */
    double det_dist;         /* metres */
    double timing_window;    /* seconds */
    int    fast;             /* 0 = reference operation order; 1 = fast mode (one rsqrt + 2 exp per sample, <= 1e-9 rel.);
                                2 = fast mode with single-precision rsqrt/exp/expm1 (optional FP32 mode, <= 1e-5 rel.) */
    double k1, k2;           /* fast mode: neg_dist_mbm*ma*width and neg_len_mbm*ma*width (0 when width <= 0) */
} alpk_prop_t;

/* event weights: generators.py:88-92 (decays), :41-46 (compton), :80-86 (inverse_primakoff) */
typedef struct {
    double days_sec;         /* days_exposure * S_PER_DAY */
    float  days_sec32;       /* float32(days_sec): Python scalar * float32 array stays float32 */
    int    days_f64;         /* exposure was an np.float64: product evaluated in f64 */
    double threshold;        /* MeV */
} alpk_event_t;

/* E[n], w[n] (f64) -> decay_w32[n], scatter_w32[n] (the reference's float32 attributes); the optional f64 outputs
 * are the weights before the float32 cast and acceptance factor (what the 1e-9 FP64 criterion is evaluated on). */
int alpk_propagate(const double* energy, const double* weight, uint64_t n, const alpk_prop_t* h_prop,
                   float* decay_w32, float* scatter_w32, double* decay_w64_or_null, double* scatter_w64_or_null,
                   void* stream);

/* decays(): event_w[n] (f64, may be NULL) and *rate = sum(event_w)  -- generators.py:88-92 */
int alpk_decays(const double* energy, const float* decay_w32, uint64_t n, const alpk_event_t* h_ev,
                double* event_w_or_null, double* rate, double* scratch, void* stream);

/* scatter channels: kind 0 = inverse Compton (generators.py:41-46, det_xs.py:89-98),
 *                   kind 1 = inverse Primakoff (generators.py:80-86, det_xs.py:37-42);
 *                   kind 2 / 3 = the same two in the fast arithmetic mode (<= 1e-9 relative: reciprocals instead of divisions).
 * h_xs: kind 0 -> {ma, ma**2, ma**4, e_thr, z*ALPHA*(g/M_E)**2, ma**4 + 2 (ma M_E)**2}
 *       kind 1 -> {ma, ma**2, (g z)**2/(2*137), r0**2, 0, 0}
 * prefactor = ntargets/det_area (applied after days*S_PER_DAY), mbm2 = METER_BY_MEV**2. */
int alpk_scatter_events(int kind, const double* energy, const float* scatter_w32, uint64_t n,
                        const double* h_xs, double prefactor, double mbm2, const alpk_event_t* h_ev,
                        double* event_w_or_null, double* rate, double* scratch, void* stream);

/* pair production a e- -> e- e+ e-: ElectronEventGenerator.pair_production (generators.py:33-39) with the tabulated
 * ALPPairProdutionCrossSection.sigma_mev (photon_xs.py:122-125): heaviside(E-2me,0)*10**interp(log10 E; left=-inf).
 * log10_xs = log10(argon_rescale*xs_dim*sigma_i/MEV2_CM2); lead = days*S_PER_DAY*(ntargets/det_area)*ge**2. */
int alpk_pair_events(const double* energy, const float* scatter_w32, uint64_t n, const double* log10_e,
                     const double* log10_xs, int n_table, double lead, double mbm2, const alpk_event_t* h_ev,
                     double* event_w_or_null, double* rate, double* scratch, void* stream);

/* ---- tabulated cross sections as functions of E (photon_xs.py:45-52 AbsCrossSection.sigma_cm2 / mu; :117-128
 * ALPPairProdutionCrossSection.sigma_cm2 / sigma_mev / mu): log-log interpolation of a table whose unit factor the caller
 * folded into log10_xs.  kind 0: 10**interp(left = 0, right = 0); kind 1: heaviside(E-2me,0)*10**interp(left = -inf). */
int alpk_table_sigma(int kind, const double* log10_e, const double* log10_xs, int n_table, const double* energy, uint64_t n,
                     double* sigma, void* stream);

/* ---- photon absorption table: photon_xs.py:48-49 (AbsCrossSection.sigma_mev) --------------------------------- */
int alpk_abs_sigma_mev(const double* log10_e, const double* log10_xs, int n_table, const double* energy, uint64_t n,
                       double* sigma_mev, void* stream);

/* ---- Primakoff production: fluxes.py:133-142, prod_xs.py:99-104 ----------------------------------------------------
 * rows already filtered by the host (E_gamma >= ma, fluxes.py:136).  pref = (g z)^2/(2*137), r02 = r0^2. */
int alpk_primakoff(const double* e_gamma, const double* phi_gamma, uint64_t n_rows,
                   const double* log10_e, const double* log10_xs, int n_table,
                   double ma, double pref, double r02, double* out_energy, double* out_flux, void* stream);
/* The same production fused with propagate (fluxes.py:43-65, 159-162) and the decay event weights + rate
 * (generators.py:88-92) in ONE pass over the rows (the deferred schedule of FluxPrimakoffIsotropic + PhotonEventGenerator.decays):
 * every output array is optional (NULL = not written); arrays and rate have the bits of alpk_primakoff + alpk_propagate +
 * alpk_decays. */
int alpk_primakoff_chain(const double* e_gamma, const double* phi_gamma, uint64_t n_rows,
                         const double* log10_e, const double* log10_xs, int n_table,
                         double ma, double pref, double r02, const alpk_prop_t* h_prop, const alpk_event_t* h_ev,
                         double* out_energy, double* out_flux, float* decay_w32, float* scatter_w32,
                         double* decay_w64, double* scatter_w64, double* event_w, double* rate, double* scratch,
                         void* stream);

/* ---- Compton production: fluxes.py:210-224, prod_xs.py:155-169 ------------------------------------------------------ */
typedef struct {
    double ma, ma2;          /* ma, ma**2 */
    double aa;               /* g**2/4/pi */
    double z;                /* target Z */
    double n_samples;        /* n_samples as a double */
    int    fast;             /* 1 = fast mode: shared reciprocal, per-row factors pre-folded (<= 1e-9 relative) */
    int    rng_rounds;       /* Philox rounds of the free-running sample stream: 10 (or 0) = Philox4x32-10, 7 = Philox4x32-7 */
} alpk_compton_t;

/* per-row table [n_rows][ALPK_COMPTON_NP]; rows already filtered by the host (s >= (M_E+ma)^2, fluxes.py:214-216) */
int alpk_compton_rows(const double* e_gamma, const double* phi_gamma, uint64_t n_rows,
                      const double* log10_e, const double* log10_xs, int n_table,
                      const alpk_compton_t* h_par, double* row_table, void* stream);

/* Per-sample kernel.  n = n_rows*n_samples samples, row-major (all samples of row 0, then row 1, ...).
 * Outputs (each may be NULL): out_energy/out_flux f64 (axion_energy/axion_flux), and -- when h_prop != NULL -- the
 * fused propagate (decay_w32/scatter_w32, optional f64 pre-cast weights) and -- when h_ev != NULL too -- the fused
 * decays() event weights and total rate.  row_ids: global row index per table row for the Philox key (NULL = 0..).
 * hist != NULL: also accumulate np.histogram(E_a, bins=hist_edges[n_edges], weights=w) into hist[n_edges-1] with
 * per-CTA privatised bins in shared memory (w = the last fused stage's weight: event weight, else decay_w32, else
 * the production weight).  Same trailing arguments for alpk_brem_samples / alpk_pairann_samples. */
int alpk_compton_samples(const double* row_table, uint64_t n_rows, uint32_t n_samples, const uint32_t* row_ids,
                         const alpk_compton_t* h_par, const double* uniforms, uint64_t seed,
                         double* out_energy, double* out_flux,
                         const alpk_prop_t* h_prop, float* decay_w32, float* scatter_w32,
                         double* decay_w64, double* scatter_w64,
                         const alpk_event_t* h_ev, double* event_w, double* rate, double* scratch,
                         const double* hist_edges, int n_edges, double* hist, void* stream);

/* ---- a -> gamma gamma decay MC: generators.py:94-128, cross_section_mc.py:510-544, 669-685 ------------------------------
 * parents: energy[n_parents] f64 and decay_w32[n_parents]; n_samples decays each, parent-major order.
 * outputs SoA: p1[4][N], p2[4][N] (E,px,py,pz planes, N = n_parents*n_samples) and w[N] (f64 holding the float32
 * event weight days*S_PER_DAY*w32/n_samples, generators.py:97,126).
 * uniforms (parity mode): [n_parents][2][n_samples] -- per parent n_samples phi then n_samples theta draws. */
int alpk_decay_parent_rows(const double* energy, const float* decay_w32, uint64_t n_parents, double ma2,
                           const alpk_event_t* h_ev, uint32_t n_samples, double* row_table, void* stream);
int alpk_decay2body(const double* row_table, uint64_t n_parents, uint32_t n_samples, const uint32_t* row_ids,
                    const double* uniforms, uint64_t seed, int fast, double* p1, double* p2, double* w, void* stream);

/* Optional FP32 mode of alpk_decay2body (north star: 1e-5 relative): float planes p1[4][N], p2[4][N], w[N] (36 B/event),
 * per-parent constants still from the FP64 row table, 24-bit Philox mantissas (event s: counter s>>1, word pair s&1,
 * stream tag kTagDecay2+16).  Free-running RNG only. */
int alpk_decay2body_f32(const double* row_table, uint64_t n_parents, uint32_t n_samples, const uint32_t* row_ids,
                        uint64_t seed, float* p1, float* p2, float* w, void* stream);

/* General 2-body decay: parents with arbitrary direction (rows {E, px, py, pz}), daughter masses m1, m2, dense 4x4
 * boost -- Decay2Body + lorentz_boost (cross_section_mc.py:496-544, 669-685) as used for parents with a polar angle
 * (generators.py:117-121) and by FluxNeutralMeson2BodyDecay (fluxes.py:1026-1031).  parent_w (may be NULL -> 1): weight
 * copied to every event of the parent.  row_table: scratch [n_parents][ALPK_PARENT_GENERAL_NP].  Outputs: p1[4][N]
 * (required), p2[4][N], w[N], theta1[N] = polar angle of daughter 1 (each may be NULL).  uniforms as alpk_decay2body. */
int alpk_decay2body_general(const double* parents_p4, const double* parent_w, uint64_t n_parents, double m1, double m2,
                            uint32_t n_samples, const uint32_t* row_ids, const double* uniforms, uint64_t seed,
                            double* row_table, double* p1, double* p2, double* w, double* theta1, void* stream);

/* ---- bremsstrahlung with track-length attenuation: fluxes.py:257-356, prod_xs.py:209-221, 313-323 ------------------------ */
typedef struct {
    double ma, ma2;
    double log10_ma;         /* np.log10(ma) */
    double ep_min;           /* max(ma, M_E)                                        fluxes.py:346 */
    double nt;               /* ntarget_area_density * HBARC**2                     fluxes.py:290,330 */
    double pref_num;         /* 2*r0**2*g**2/4/pi, r0 = ALPHA/M_E                   prod_xs.py:212,217 */
    double zl;               /* z**2*ln_el + z*ln_inel                              prod_xs.py:215-219 */
    double zz;               /* z**2 + z */
    double pref_vec;         /* chi*(4*ALPHA**2*g**2)/(4*pi)                        prod_xs.py:317-322 */
    double ln10;             /* np.log(10) */
    double n_samples;
    double t_arg;            /* 4*max_track_length/3                                fluxes.py:266 */
    int vector;              /* boson_type == "vector" */
    int use_track_length;    /* simulate(use_track_length=...) */
    int fast;                /* 1: fast arithmetic mode of the per-sample production (<= 1e-9 relative); must equal the mode of a
                                fused alpk_prop_t so that every schedule produces the same bits */
    int rng_rounds;          /* as in alpk_compton_t (the per-sample E_a stream; the per-row draw stays Philox4x32-10) */
} alpk_brem_t;

/* Row prelude: one thread per electron row (rows already filtered by the host: E0 >= ep_min in track-length mode).
 * u_rows (parity mode) / Philox (seed, row_ids): the per-row scalar draw of fluxes.py:350.  The e-/e+ dN/dE tables
 * (np.interp, left=right=0) and the 32-point Gauss-Legendre rule gl_x/gl_w[ALPK_GL_POINTS] for the track-length
 * integral (fmath.py:33-36) are device arrays.  Writes row_table[n_rows][ALPK_BREM_NP] and emit[n_rows]
 * (1 = ea_max > ma, the row produces n_samples samples; fluxes.py:322-324). */
int alpk_brem_rows(const double* e0, const double* w0, uint64_t n_rows, const uint32_t* row_ids,
                   const double* u_rows, uint64_t seed,
                   const double* ele_e, const double* ele_f, int n_ele, const double* pos_e, const double* pos_f, int n_pos,
                   const double* gl_x, const double* gl_w, const alpk_brem_t* h_par,
                   double* row_table, uint8_t* emit, void* stream);

/* Per-sample kernel over the EMITTING rows (row table compacted by the caller); same fused tail as
 * alpk_compton_samples.  uniforms (parity mode): [n_rows][n_samples]. */
int alpk_brem_samples(const double* row_table, uint64_t n_rows, uint32_t n_samples, const uint32_t* row_ids,
                      const alpk_brem_t* h_par, const double* uniforms, uint64_t seed,
                      double* out_energy, double* out_flux,
                      const alpk_prop_t* h_prop, float* decay_w32, float* scatter_w32,
                      double* decay_w64, double* scatter_w64,
                      const alpk_event_t* h_ev, double* event_w, double* rate, double* scratch,
                         const double* hist_edges, int n_edges, double* hist, void* stream);

/* ---- resonant production: fluxes.py:421-445.  *out_sum = sum_i Phi_p(E_i) * track_length_prob(E_i, E_res, t_i) over
 * n_samples draws (E_i, t_i) ~ U(E_res, E_max) x U(0, max_t); the caller applies the scalar prefactors (:441-442).
 * uniforms (parity mode): [2][n_samples] -- the E draws then the t draws (:437-438). */
int alpk_resonance_sum(const double* pos_e, const double* pos_f, int n_pos, double e_res, double e_max, double max_t,
                       uint64_t n_samples, const double* uniforms, uint64_t seed, int fast, double* out_sum, double* scratch,
                       void* stream);   /* fast: 1 = fast arithmetic mode (<= 1e-9 relative: exp-table pow, polynomial 1/Gamma) */

/* ---- associated production e+ e- -> a gamma: fluxes.py:503-518, cross_section_mc.py:192-240, matrix_element.py:717-734 -- */
typedef struct {
    double ma, ma2;
    double me2, me4;         /* M_E**2, M_E**4 */
    double pref;             /* (pi*ALPHA)*coupling_product**2, coupling_product = 1 */
    double wconst;           /* reserved */
    double z, ge2, ntad, hbarc2;   /* applied left to right after the positron weight, fluxes.py:518 */
    double n_samples;
    int fast;                /* as in alpk_brem_t */
    int rng_rounds;          /* as in alpk_compton_t */
} alpk_pairann_t;

/* rows already filtered by the host (E+ >= max((ma^2-me^2)/(2 me), me), fluxes.py:507) */
int alpk_pairann_rows(const double* e_pos, const double* w_pos, uint64_t n_rows, const alpk_pairann_t* h_par,
                      double* row_table, void* stream);
/* uniforms (parity mode): [n_rows][2][n_samples] -- phi draws then theta draws per row (cross_section_mc.py:204-205);
 * only the theta draws enter the lab energy and weight of a beam along z. */
int alpk_pairann_samples(const double* row_table, uint64_t n_rows, uint32_t n_samples, const uint32_t* row_ids,
                         const alpk_pairann_t* h_par, const double* uniforms, uint64_t seed,
                         double* out_energy, double* out_flux,
                         const alpk_prop_t* h_prop, float* decay_w32, float* scatter_w32,
                         double* decay_w64, double* scatter_w64,
                         const alpk_event_t* h_ev, double* event_w, double* rate, double* scratch,
                         const double* hist_edges, int n_edges, double* hist, void* stream);

/* ---- propagate_iso_vol_int: fluxes.py:67-89 with DetectorGeometry.integrate materials.py:98-124 ---------------------------
 * Per sample the Monte-Carlo volume integral of (1/(hbar c v tau)) exp(-l/(hbar c v tau))/(4 pi l^2) over the
 * detector points.  Host-prepared per point: nxm = -l/METER_BY_MEV, bden = 4*pi*l**2, jac = 1 (box), u1**2*sin(u2)
 * (sphere), u1 (cylinder).  volume_over_n: the geometry's volume factor / n_points; inv_mbm = 1/METER_BY_MEV.
 * decay_w = float32(rescale*w*integral) (no acceptance factor), scatter_w = float32(rescale*w*surv).
 * partials: scratch of alpk_vol_int_chunks(n, n_points) * n doubles. */
uint32_t alpk_vol_int_chunks(uint64_t n, uint32_t n_points);
int alpk_vol_int(const double* energy, const double* weight, uint64_t n, const alpk_prop_t* h_prop,
                 const double* nxm, const double* bden, const double* jac, uint32_t n_points,
                 double volume_over_n, double inv_mbm, float* decay_w32, float* scatter_w32,
                 double* decay_w64_or_null, double* scatter_w64_or_null, double* partials, void* stream);

/* ---- small flux sources ---------------------------------------------------------------------------
 * nuclear de-excitation lines: FluxNuclearIsotropic.simulate/br (fluxes.py:571-590).  rates4[n][4] = {E, rate, beta, eta}
 * (rows with E > ma, filtered by the host); c0 = (j/(j+1))/(1+delta**2)/pi/ALPHA; expo = 2j+1. */
int alpk_nuclear(const double* rates4, uint64_t n_rows, double ma2, double c0, double expo, double gann0, double gann1,
                 double* out_energy, double* out_flux, void* stream);
/* pi0 -> gamma a in flight: FluxPi0Isotropic.simulate_flux (fluxes.py:941-968).  One cos(theta*) draw per pi0 momentum
 * (uniforms[n] in parity mode); keep[i] = 1 if the sample passes the energy / angle cuts (the caller compacts). */
int alpk_pi0_flux(const double* momenta, uint64_t n, const double* uniforms, uint64_t seed, double mm2, double ma2,
                  double p_cm, double e1_cm, double energy_cut, double angle_cut, double weight,
                  double* out_energy, double* out_flux, uint8_t* keep, void* stream);

/* e+ e- -> gamma a through a virtual photon: FluxPairAnnihilationGamma.simulate (fluxes.py:1090-1134) with
 * epem_to_alp_photon_dsigma_de (prod_xs.py:109-129).  bins = positron-table bin centres above threshold (host-filtered),
 * n_samples (depth, smeared E+, E_a) draws each.  Parity mode: u_bins[n_bins][2][n_samples] then u_alp[n_bins*n_samples]. */
typedef struct {
    double ma, ma2, ma4;
    double me2, me4;
    double ep_min;           /* max((ma**2 - M_E**2)/(2*M_E), M_E) */
    double nt;               /* ntarget_area_density * HBARC**2 */
    double zc;               /* z*4*pi*ALPHA*g**2 */
    double n_samples;
    double max_t;            /* 5.0 */
    int fast;                /* 1: fast arithmetic mode of the track-length factor (<= 1e-9 relative) */
} alpk_pairgamma_t;
int alpk_pairgamma(const double* centers, const double* widths, uint64_t n_bins, uint32_t n_samples,
                   const uint32_t* bin_ids, const double* pos_e, const double* pos_f, int n_pos,
                   const alpk_pairgamma_t* h_par, const double* u_bins, const double* u_alp, uint64_t seed,
                   double* out_energy, double* out_flux, void* stream);

/* planes[ncomp][n] -> rows[n][ncomp] (ncomp 3 or 4): the per-event layout of LorentzVector.pmu / Vector3.vec
 * (cross_section_mc.py:14-143) for a single contiguous device -> host copy of all events.  rows 16-byte aligned. */
int alpk_planes_to_rows(const double* planes, uint64_t n, int ncomp, double* rows, void* stream);

/* out[i] = factor * x[i]  (float64 weights of FluxPi0Isotropic.propagate, fluxes.py:985-991) */
int alpk_scale(const double* x, uint64_t n, double factor, double* out, void* stream);

/* ---- weighted histogram: np.histogram(x, bins=edges, weights=w) (half-open bins, last bin closed) ------------------ */
int alpk_hist(const double* x, const double* w64_or_null, const float* w32_or_null, uint64_t n,
              const double* edges, int n_edges, double* hist /* [n_edges-1], accumulated into */, void* stream);

/* ---- batched (m_a, g) scan: README.md:199-226 / examples/ex4_na64.py:30-40 pattern (one simulate() per m_a, one
 * propagate(new_coupling=g) + decays() per point, fluxes.py:153-162,235-245; generators.py:88-92) in one launch grid.
 * process 0 = Compton  (c0 = g0**2/4/pi, z = target Z, n_samples per row),
 * process 1 = Primakoff (c0 = (g0 z)**2/(2*137), c1 = r0**2; one sample per row).
 * e_gamma/phi_gamma: the FULL flux table (rows failing a mass's threshold contribute nothing for that mass);
 * row_offset: global index of row 0 (Philox key) when the table is a shard.
 * ma_grid/ma2_grid[n_ma] (ma, ma**2); width[n_ma][n_g] decay widths; rescale[n_g] = (g/g0)**2.
 * h_geom: detector geometry / acceptance / fast flag of alpk_prop_t (ma, width, rescale fields ignored).
 * rng_rounds: Philox rounds of the Compton sample stream (as alpk_compton_t: must match the per-point API's to reproduce it).
 * tables: scratch [n_ma][n_rows][ALPK_COMPTON_NP or 2]; partials: scratch [n_ma][partial_tiles][n_g] with
 * 1 <= partial_tiles <= alpk_scan_tiles(): the tile axis (1024 samples per tile) is walked in chunks of partial_tiles, so
 * the scratch stays bounded for any sample count; the sums run in tile order whatever the chunking.
 * out_rates[n_ma][n_g] = decays(days, threshold) of every grid point. */
uint32_t alpk_scan_tiles(uint32_t n_rows, uint32_t n_samples);
int alpk_scan(int process, const double* e_gamma, const double* phi_gamma, uint32_t n_rows, uint32_t row_offset,
              const double* log10_e, const double* log10_xs, int n_table,
              const double* ma_grid, const double* ma2_grid, int n_ma,
              const double* width, const double* rescale, int n_g,
              double c0, double c1, double z, uint32_t n_samples, uint64_t seed, int rng_rounds,
              const alpk_prop_t* h_geom, const alpk_event_t* h_ev,
              double* tables, double* partials, uint32_t partial_tiles, double* out_rates, void* stream);

/* FP64 issue-rate microbenchmark: `blocks` CTAs x 256 threads x 8 independent DFMA chains x iters (bench.py uses it
 * to measure the FP64 roofline denominator on the box).  out: >= blocks*256 doubles (never written in practice). */
int alpk_fp64_peak(uint64_t iters, int blocks, double* out, void* stream);
/* FP64 exp throughput: `blocks` CTAs x 256 threads x 4*iters evaluations; kind 0 = library exp, 1 = the fast-mode table exp */
int alpk_exp_peak(int kind, uint64_t iters, int blocks, double* out, void* stream);

/* ---- peer windows: the path's one exchange step over NVLink peer memory ------------------------------------------------
 * The reference is single-process; its only cross-sample steps are np.sum / np.histogram of the per-sample weights
 * (generators.py:27-30, 48-52) and the user's loop over a (m_a, g) grid (README.md:199-226).  When those are sharded over
 * GPUs (one process per GPU), the ranks' partial [rates || histogram bins] or their rows of a rate map are merged by ONE
 * single-CTA kernel per rank that pushes its values into every rank's window with NVLink stores, signals, waits for the
 * other ranks' signals and combines the slots in rank order (alpk_peer.cu) -- no library collective, no host round trip,
 * replayable from a captured CUDA graph.  Every rank gets the same bits, independent of arrival order.
 *
 * A window is created by its owner (cudaMalloc + cudaIpcGetMemHandle), the 64-byte handle is sent to the other processes
 * of the node by any means (alplib_b200.peer uses torch.distributed's object all-gather), and they map it with
 * alpk_peer_window_open (cudaIpcOpenMemHandle; peer access is enabled on demand).  All ranks must issue the same sequence
 * of merges on a window, each rank on ONE stream.  capacity = doubles per payload half: an all-reduce needs
 * n_ranks * n_total <= capacity, an all-gather n_total_rows * row_len <= capacity.  timeout_ns = 0 -> 10 s; a merge that
 * times out (a rank died or never launched) writes NaN to `out` and sets the status word -- it does not hang the GPU. */
#define ALPK_PEER_MAX_RANKS 16
#define ALPK_PEER_MAX_PARTS 8
#define ALPK_PEER_HANDLE_BYTES 64
typedef struct {
    void* windows[ALPK_PEER_MAX_RANKS];   /* device pointers of every rank's window as mapped in THIS process (own included) */
    int rank, n_ranks;
    uint64_t capacity;                    /* doubles per payload half (the same on every rank) */
} alpk_peer_t;
uint64_t alpk_peer_window_bytes(uint64_t capacity_doubles);
int alpk_peer_window_create(uint64_t capacity_doubles, void** h_window, unsigned char* h_handle);
int alpk_peer_window_open(const unsigned char* h_handle, void** h_window);
int alpk_peer_window_close(void* d_window);        /* a window mapped with _open */
int alpk_peer_window_destroy(void* d_window);      /* a window made with _create */
/* number of merges completed on this rank's window and its status (0 ok, 1 a merge timed out); synchronises `stream` */
int alpk_peer_status(const alpk_peer_t* h_peer, uint64_t* h_epoch, int* h_status, void* stream);
/* out[n_total] = sum over ranks of the concatenation of the n_parts segments (h_parts[i]: device pointer, h_part_len[i]
 * doubles), added in rank order */
int alpk_peer_all_reduce_sum(const alpk_peer_t* h_peer, const double* const* h_parts, const uint32_t* h_part_len, int n_parts,
                             double* out, uint64_t timeout_ns, void* stream);
/* out[n_total_rows][row_len]: local row i of local[n_local_rows][row_len] is global row row_start + i * row_stride */
int alpk_peer_all_gather_rows(const alpk_peer_t* h_peer, const double* local, uint32_t n_local_rows, uint32_t row_len,
                              uint32_t row_start, uint32_t row_stride, uint32_t n_total_rows, double* out,
                              uint64_t timeout_ns, void* stream);

/* sum of a f64 array (deterministic two-stage) */
int alpk_sum(const double* x, uint64_t n, double* out, double* scratch, void* stream);

/* ---- charged-meson 3-body decay fluxes M -> l nu X: FluxChargedMeson3BodyIsotropic (fluxes.py:811-922) and
 * FluxChargedMeson3BodyDecay (fluxes.py:605-808), with M2Meson3BodyDecay (matrix_element.py:419-518) integrated over the
 * Dalitz variable m_23^2 by the 21-point Gauss-Kronrod rule QUADPACK applies for scipy.integrate.quad (fluxes.py:639-661).
 * model: 0 scalar_ib1, 1 pseudoscalar_ib1, 2 vector_ib1, 3 scalar_ib2, 4 vector_ib2, 5 vector_ib9, 6 sd,
 * 7 vector_contact, 8 any other string (zero matrix element). */
typedef struct {
    double mm, ma;            /* meson mass, boson mass */
    double fM, ckm;           /* decay constant, CKM element */
    double total_width;       /* meson total width [MeV] */
    double c0;                /* contact-model parameter */
    double coupling;
    double a, b, d;           /* vector_ib9 free parameters (abd) */
    double pref;              /* (2*mm)/(32*power(2*pi*mm, 3)) */
    double n_samples;         /* n_samples as a double (divisor) */
    int model;
} alpk_meson3_t;
#define ALPK_MESON3_NP 13
/* rows: one per (meson, allowed lepton) -- row_meson[r] indexes `mesons`, row_ml[r] is the lepton mass.  mode 0:
 * mesons[n][2] = {p, weight}; mode 1: mesons[n][3] = {p, theta, weight} and one exponential decay position per meson
 * (u_pos[n_mesons] in parity mode, else Philox keyed by meson index + meson_offset).  row_table: [n_rows][ALPK_MESON3_NP];
 * entry 12 is the row's live flag (decay position <= 52 m, fluxes.py:671). */
int alpk_meson3_rows(int mode, const double* mesons, uint64_t n_mesons, const uint32_t* row_meson, const double* row_ml,
                     uint64_t n_rows, const alpk_meson3_t* h_par, double det_dist, double det_length, const double* u_pos,
                     uint64_t seed, uint32_t meson_offset, double* row_table, void* stream);
/* samples: n_samples per row.  Parity mode: uniforms[block][2 or 3][n_samples] (energies, cosines[, azimuths]) with
 * block_of_row[r] the block of live row r.  mode 0 writes out_energy / out_flux (the rest may be NULL); mode 1 writes all
 * arrays and keep[i] = 1 for samples of live rows that pass the solid-angle acceptance (or all, if the cut is off). */
int alpk_meson3_samples(int mode, const double* row_table, uint64_t n_rows, uint32_t n_samples, const alpk_meson3_t* h_par,
                        double energy_cut, int cut_on_solid_angle, const double* uniforms, const int32_t* block_of_row,
                        uint64_t seed, uint32_t row_offset, double* out_energy, double* out_flux, double* out_cos,
                        double* out_theta, double* out_solid_angle, double* out_decay_pos, uint8_t* keep, void* stream);
/* dGammadEa(E_a, ml) (fluxes.py:639-661 / 841-863) on an array of energies */
int alpk_meson3_dgamma(const double* ea, uint64_t n, double lepton_mass, const alpk_meson3_t* h_par, double* out, void* stream);
/* FluxAxionMesonMixing.simulate (fluxes.py:1202-1224): p4[n][4] = {E, px, py, pz}; writes E, the polar angle and the
 * (constant) rate per meson, keep[i] = 1 for mesons with E >= m_a inside the half opening angle det_sa. */
int alpk_meson_mixing(const double* p4, uint64_t n, double ma, double det_sa, double rate, double* out_energy,
                      double* out_angle, double* out_flux, uint8_t* keep, void* stream);
/* M2Meson3BodyDecay.__call__(m122, m232) (matrix_element.py:443-518) on an array of m232 */
int alpk_meson3_m2(double m122, const double* m232, uint64_t n, double lepton_mass, const alpk_meson3_t* h_par, double* out,
                   void* stream);

/* ---- generic 2 -> 2 scattering MC: cross_section_mc.py:148-275 (Scatter2to2MC.scatter_sim, dsigma_dt,
 * get_cosine_lab_weights, get_e3_lab_weights, get_e4_lab_weights) with the squared matrix elements of
 * matrix_element.py:629-734 (M2Compton, M2InverseCompton, M2InversePrimakoff, M2Primakoff, M2AssociatedProduction) and the
 * atomic elastic form factor form_factors.py:54-64; usage README.md:144-158. ------------------------------------------ */
enum { ALPK_M2_INVERSE_PRIMAKOFF = 0, ALPK_M2_PRIMAKOFF = 1, ALPK_M2_INVERSE_COMPTON = 2, ALPK_M2_COMPTON = 3,
       ALPK_M2_ASSOCIATED_PRODUCTION = 4 };

typedef struct {
    int    model;            /* ALPK_M2_* */
    double ma2, ma4;         /* ma**2, ma**4 */
    double mN2, mN4;         /* mN**2, mN**4 (Primakoff-type) */
    double z;                /* atomic number */
    double ff_a2;            /* a**2 of AtomicElasticFF (form_factors.py:63) */
    double pref;             /* prefactor incl. coupling_product**2: z*4*pi*ALPHA*cp^2 | 2*cp^2 | cp^2 | pi*ALPHA*cp^2 */
    double me2, me4;         /* M_E**2, M_E**4 */
    double tmax;             /* associated production: soft-photon cut (matrix_element.py:729-731) */
} alpk_m2_t;

/* out[i] = M2(s[i*s_stride], t[i]) (den == 0) or M2/den (dsigma_dt, cross_section_mc.py:184-185 with
 * den = 16*pi*(s-(m1+m2)^2)*(s-(m1-m2)^2)); s_stride 0 broadcasts one s. */
int alpk_m2_eval(const alpk_m2_t* h_m2, const double* s, uint64_t s_stride, const double* t, uint64_t n, double den,
                 double* out, void* stream);

/* AtomicElasticFF.__call__ (form_factors.py:61-64): out[i] = (z*(q^2 a2)/(1 + q^2 a2))^2 */
int alpk_atomic_ff2(const double* q, uint64_t n, double z, double a2, double* out, void* stream);

typedef struct {
    alpk_m2_t m2;
    double s;                /* (lv_p1 + lv_p2).mass2() */
    double msum;             /* m1**2 + m3**2 */
    double pp, ee;           /* p1_cm*p3_cm, e1_cm*e3_cm */
    double p3_cm, e3_cm;
    double vol;              /* 4*p1_cm*p3_cm/n_samples (linear sampling) */
    double p1_cm, n_samples; /* log sampling: exp(logu)*2*p1_cm*p3_cm/n_samples */
    double den;              /* 16*pi*(s-(m1+m2)^2)*(s-(m1-m2)^2) */
    double mat[16];          /* lorentz_boost matrix for -v_in, row-major (identity if beta == 0) */
    double p12[4];           /* lv_p1 + lv_p2 */
    int    log_sampling;
} alpk_scatter2_t;

/* One scatter_sim(): n_samples draws (phi, theta).  uniforms (parity mode): [2][n] -- the phi block then the theta (or
 * logu) block, the reference's draw order; NULL: Philox keyed (seed, tag 8, stream_id, sample).  Outputs (each may be
 * NULL): t[n], weights[n] = dsigma_dcos_cm_wgts, component planes p3_cm[4][n], p3_lab[4][n], p4_lab[4][n],
 * cosine_lab[n] = p3_lab.cosine(), 3-velocity planes v3_cm[3][n], v3_lab[3][n], v4_lab[3][n]. */
int alpk_scatter2to2(const alpk_scatter2_t* h_par, uint64_t n_samples, const double* uniforms, uint64_t seed,
                     uint32_t stream_id, double* t_out, double* weights, double* p3_cm, double* p3_lab, double* p4_lab,
                     double* cosine_lab, double* v3_cm, double* v3_lab, double* v4_lab, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALPK_H */
