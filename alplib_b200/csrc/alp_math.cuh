// Per-sample / per-bin arithmetic of the alplib Monte-Carlo hot path, written once and used by every kernel.
//
// Everything here is `ALP_HD` (host + device) so that the exact same expressions can be exercised on the CPU by
// tests/hostcheck (unit tests only -- the product path is the CUDA kernels, there is no CPU fallback).
//
// FP64 op order follows the reference expression by expression (file:line cited per function, relative to the
// reference repository root).  The library is compiled with -fmad=false so that no a*b+c is contracted into an
// FMA: several expressions are cancellation-dominated (parent mass from the 4-vector, gamma from 1/sqrt(1-b^2),
// the Compton x variable) and a fused rounding there moves results by far more than the 1e-9 parity tolerance.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ALP_HD __host__ __device__ __forceinline__
#else
#define ALP_HD inline
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
#endif

namespace alp {

// ---- constants.py:14-66 -------------------------------------------------------------------------------------
constexpr double kAlpha = 1.0 / 137.0;                     // constants.py:14  ALPHA = 1/137
constexpr double kME = 0.511;                              // constants.py:32
constexpr double kMeterByMeV = 6.58212e-22 * 2.998e8;      // constants.py:60
constexpr double kMeV2Cm2 = (kMeterByMeV * 100) * (kMeterByMeV * 100);   // constants.py:65
constexpr double kHbarC = 1.97236e-11;                     // constants.py:22
constexpr double kAvogadro = 6.022e23;                     // constants.py:29
constexpr double kCLight = 2.99792458e10;                  // constants.py:20
constexpr double kPi = 3.141592653589793;                  // numpy.pi
constexpr double kTwoPi = 2 * 3.141592653589793;           // 2*pi as evaluated by Python

// ---- counter-based RNG: Philox4x32-10 (Salmon et al., SC'11) ------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

ALP_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    const uint64_t p = (uint64_t)a * (uint64_t)b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
}

ALP_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c0, hi0, lo0);
        mulhilo32(M1, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// The key schedule (k0 + r W0, k1 + r W1) depends only on the seed: kernels take it precomputed (kernel parameter ->
// constant bank), which turns each round's key bump + xor into one LOP3 with a constant operand.
struct PhiloxKeys { uint32_t k[20]; };

ALP_HD PhiloxKeys philox_keys(uint64_t seed) {
    PhiloxKeys rk;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { rk.k[2 * r] = k0; rk.k[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return rk;
}

// R rounds with the precomputed key schedule.  R = 10 is Philox4x32-10 (the Random123 / cuRAND default); R = 7 is
// Philox4x32-7, the fewest rounds Salmon et al. (SC'11, table 2) found Crush-resistant -- the fast-mode choice for the
// per-sample production streams, where the generator is 1/5 of the chain kernel's issue slots.
template <int R>
ALP_HD Philox4 philox4x32_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& rk) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo32(M0, c0, hi0, lo0);
        mulhilo32(M1, c2, hi1, lo1);
        const uint32_t n0 = hi1 ^ c1 ^ rk.k[2 * r];
        const uint32_t n2 = hi0 ^ c3 ^ rk.k[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return Philox4{c0, c1, c2, c3};
}

// 53-bit uniform in [0,1) from two 32-bit words (same construction as MT19937 genrand_res53 / numpy legacy
// random_sample, so the free-running stream has the reference's granularity).
ALP_HD double u53(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    // The same value built in the mantissa fields (no int->double conversions, every step exact):
    //   A = 2^25 + (a>>5) 2^-27,  B = 2^-1 + (b>>6) 2^-53,  u = (A - (2^25 + 2^-1)) + B
    const double A = __hiloint2double(0x41800000, (int)(a >> 5));
    const double B = __hiloint2double(0x3FE00000, (int)(b >> 6));
    return (A - 33554432.5) + B;
#else
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
#endif
}

// ---- branch-free reciprocal / reciprocal square root for the fast mode --------------------------------------------
// Hardware seed (MUFU.RCP64H / MUFU.RSQ64H, ~2^-22 relative) + one cubically convergent correction: ~1 ulp for normal,
// finite, non-zero arguments; no special-case branch (callers guard the argument range), so several of them interleave.
#if !defined(__CUDA_ARCH__) && defined(ALP_EMULATE_APPROX_SEED)
// tests/hostcheck only: the hardware seeds (MUFU.RCP64H / RSQ64H, ~2^-22 relative) emulated by the exact result cut to 21
// mantissa bits, so that the CPU test tier runs the DEVICE correction steps on a seed at least as bad as the hardware's.
static inline double alp_seed21(double v) {
    uint64_t b;
    __builtin_memcpy(&b, &v, 8);
    b &= ~((1ull << 31) - 1ull);
    __builtin_memcpy(&v, &b, 8);
    return v;
}
#define ALP_HOST_SEEDS 1
#else
#define ALP_HOST_SEEDS 0
#endif

ALP_HD double rcp_fast(double a) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    e = fma(e, e, e);
    return fma(y, e, y);
#elif ALP_HOST_SEEDS
    const double y = alp_seed21(1.0 / a);
    double e = fma(-a, y, 1.0);
    e = fma(e, e, e);
    return fma(y, e, y);
#else
    return 1.0 / a;
#endif
}

ALP_HD double rsqrt_fast(double d) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = fma(-d, y * y, 1.0);
    const double p = fma(h, 0.375, 0.5);
    return fma(p, y * h, y);
#elif ALP_HOST_SEEDS
    const double y = alp_seed21(1.0 / sqrt(d));
    const double h = fma(-d, y * y, 1.0);
    const double p = fma(h, 0.375, 0.5);
    return fma(p, y * h, y);
#else
    return 1.0 / sqrt(d);
#endif
}

// One-step variants for quantities that only need the fast mode's 1e-9: seed (2^-22) + ONE quadratically convergent
// correction: ~6e-14 / ~9e-14 relative, one FP64 instruction fewer each (2 / 4 instead of 3 / 5).
ALP_HD double rcp_fast1(double a) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return fma(y, fma(-a, y, 1.0), y);
#elif ALP_HOST_SEEDS
    const double y = alp_seed21(1.0 / a);
    return fma(y, fma(-a, y, 1.0), y);
#else
    return 1.0 / a;
#endif
}

ALP_HD double rsqrt_fast1(double d) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double h = fma(-d, y * y, 1.0);
    return fma(y * 0.5, h, y);
#elif ALP_HOST_SEEDS
    const double y = alp_seed21(1.0 / sqrt(d));
    const double h = fma(-d, y * y, 1.0);
    return fma(y * 0.5, h, y);
#else
    return 1.0 / sqrt(d);
#endif
}

// Stream tags: 4th counter word.  A variate is a pure function of (seed, tag, row, sample) -- independent of the
// launch configuration and of how rows are sharded over GPUs.
enum StreamTag : uint32_t {
    kTagComptonEa = 1, kTagBremRow = 2, kTagBremEa = 3, kTagResonance = 4, kTagPairAnn = 5, kTagDecay2 = 6,
    kTagDecayParentPhi = 7
};

// One uniform per sample: sample s uses Philox counter s>>1 and word pair s&1.
// rounds: 10 (default) or 7 (see philox4x32_rk).
ALP_HD double uniform_1(uint64_t seed, uint32_t tag, uint32_t row, uint64_t s, int rounds = 10) {
    const uint64_t c = s >> 1;
    Philox4 r;
    if (rounds == 7) r = philox4x32_rk<7>((uint32_t)c, (uint32_t)(c >> 32), row, tag, philox_keys(seed));
    else r = philox4x32_10((uint32_t)c, (uint32_t)(c >> 32), row, tag, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (s & 1) ? u53(r.z, r.w) : u53(r.x, r.y);
}

// Two uniforms per sample: sample s uses Philox counter s; (x,y) -> first variate, (z,w) -> second.
ALP_HD void uniform_2(uint64_t seed, uint32_t tag, uint32_t row, uint64_t s, double& u1, double& u2) {
    const Philox4 r = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), row, tag, (uint32_t)seed, (uint32_t)(seed >> 32));
    u1 = u53(r.x, r.y);
    u2 = u53(r.z, r.w);
}

// Reduction / polynomial constants of exp_fast live in constant memory on the device so that each step is a single
// DFMA with a constant operand (64-bit immediates would cost two extra moves per step).  Table and constants are generated
// by tools/gen_exp_table.py (80-digit arithmetic): 2^11 = 2048 entries (16 KB; the sample kernels keep a copy in shared
// memory), which makes |r| <= ln2/4096 = 1.7e-4 small enough for a degree-3 polynomial (truncation r^4/24 = 3.4e-17).
#include "alp_exp_table.inc"
constexpr int kExpTabBits = ALP_EXP_TAB_BITS;
constexpr int kExpTabSize = 1 << kExpTabBits;
#if defined(__CUDACC__)
static __constant__ double kExpCoefDev[8] = ALP_EXP_COEFS;
static __constant__ double kExp10CoefDev[4] = ALP_EXP10_COEFS;
static __device__ const double kExpTabDev[kExpTabSize] = ALP_EXP_TABLE;
#endif
static const double kExpCoefHost[8] = ALP_EXP_COEFS;
static const double kExp10CoefHost[4] = ALP_EXP10_COEFS;
static const double kExpTabHost[kExpTabSize] = ALP_EXP_TABLE;
#if defined(__CUDA_ARCH__)
#define ALP_EXPC(i) kExpCoefDev[i]
#define ALP_EXP10C(i) kExp10CoefDev[i]
#else
#define ALP_EXPC(i) kExpCoefHost[i]
#define ALP_EXP10C(i) kExp10CoefHost[i]
#endif

// ---- exp for the fast mode ---------------------------------------------------------------------------------------
// Table-driven: x = (N n + j) ln2/N + r with N = 2048, |r| <= ln2/4096 (Cody-Waite, two-part ln2/N), exp(x) = 2^n 2^(j/N) exp(r);
// 2^(j/N) from the table, exp(r)-1 = r + r^2 (1/2 + r/6) (truncation 3.4e-17 relative), 2^n applied to the exponent field.
// 8 FP64 instructions, dependency depth 7.  Within 1 ulp of glibc on [-708, 708] (tests/test_hostmath.py measures it);
// anything outside (including NaN) takes the library exp.  Branch-free, so several calls interleave.
ALP_HD double exp_table_finish(double kd, double r, const double* tab) {
    const double r2 = r * r;
    const double q = fma(ALP_EXPC(4), r, ALP_EXPC(3));
    const double t = fma(q, r2, r);                                 // exp(r) - 1
#if defined(__CUDA_ARCH__)
    const int ki = __double2loint(kd);                              // N n + j, two's complement
    double tv;
    if (tab) {
        // shared-memory copy: address = base + 8 (ki & (N-1)) as one LOP3 + one IMAD (the compiler's own sequence is three)
        unsigned addr;
        asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(ki & (kExpTabSize - 1)), "r"((unsigned)__cvta_generic_to_shared(tab)));
        asm("ld.shared.f64 %0, [%1];" : "=d"(tv) : "r"(addr));
    } else {
        tv = __ldg(&kExpTabDev[ki & (kExpTabSize - 1)]);
    }
    const double S = __hiloint2double(__double2hiint(tv) + (ki << (20 - kExpTabBits)), __double2loint(tv));   // 2^n 2^(j/N)
    return fma(S, t, S);
#else
    int64_t bits, tb;
    __builtin_memcpy(&tb, &kd, 8);
    const int32_t ki = (int32_t)(uint32_t)tb;
    const double tv = tab ? tab[ki & (kExpTabSize - 1)] : kExpTabHost[ki & (kExpTabSize - 1)];
    __builtin_memcpy(&bits, &tv, 8);
    bits += (int64_t)((uint64_t)(int64_t)ki << (52 - kExpTabBits));
    double S;
    __builtin_memcpy(&S, &bits, 8);
    return fma(S, t, S);
#endif
}

// `tab`: a copy of the table in shared memory (the chain / scan / volume-integral kernels); default: the global copy through L1
ALP_HD double exp_fast_core(double x, const double* tab = nullptr) {          // requires |x| <= 708
    const double kShift = 6755399441055744.0;                       // 1.5 * 2^52: rounds to nearest integer
    const double kd = fma(x, ALP_EXPC(0), kShift);                  // N x / ln2
    const double fn = kd - kShift;
    double r = fma(fn, ALP_EXPC(1), x);
    r = fma(fn, ALP_EXPC(2), r);
    return exp_table_finish(kd, r, tab);
}

// The two exponentials of the straight-line propagate (fast mode, 1e-9 bar), each cut to what its use needs:
//   exp_surv: the survival factor exp(k1/p_a) enters the weights as a plain factor -> 1e-12 relative is ample: one-constant
//             reduction (r off by <= |n| 4e-20 <= 8e-14) and exp(r) - 1 = r + r^2/2 (truncation r^3/6 <= 8e-13): 6 FP64 instructions;
//   exp_dec:  the decay probability is 1 - exp(k2/p_a), which cancels for long-lived ALPs, so the exponential keeps its full
//             polynomial (absolute error ~1e-16 where it matters: small |x| reduces with n = 0, i.e. exactly); only the low
//             Cody-Waite term goes, whose contribution |n| 4e-20 is relevant only where exp(x) << 1 and nothing cancels: 7.
// (exp_fast_core: 8.)  Both require |x| <= 708 like exp_fast_core.
constexpr double kNegLn2OverN = -0x1.62e42fefa39efp-1 / (double)(1 << ALP_EXP_TAB_BITS);     // -ln2/N (N a power of two: exact scaling)

ALP_HD double exp_table_scale(double kd, const double* tab) {             // 2^n 2^(j/N) of exp_table_finish
#if defined(__CUDA_ARCH__)
    const int ki = __double2loint(kd);
    double tv;
    if (tab) {
        unsigned addr;
        asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(ki & (kExpTabSize - 1)), "r"((unsigned)__cvta_generic_to_shared(tab)));
        asm("ld.shared.f64 %0, [%1];" : "=d"(tv) : "r"(addr));
    } else {
        tv = __ldg(&kExpTabDev[ki & (kExpTabSize - 1)]);
    }
    return __hiloint2double(__double2hiint(tv) + (ki << (20 - kExpTabBits)), __double2loint(tv));
#else
    int64_t bits, tb;
    __builtin_memcpy(&tb, &kd, 8);
    const int32_t ki = (int32_t)(uint32_t)tb;
    const double tv = tab ? tab[ki & (kExpTabSize - 1)] : kExpTabHost[ki & (kExpTabSize - 1)];
    __builtin_memcpy(&bits, &tv, 8);
    bits += (int64_t)((uint64_t)(int64_t)ki << (52 - kExpTabBits));
    double S;
    __builtin_memcpy(&S, &bits, 8);
    return S;
#endif
}

ALP_HD double exp_surv(double x, const double* tab = nullptr) {
    const double kShift = 6755399441055744.0;
    const double kd = fma(x, ALP_EXPC(0), kShift);
    const double r = fma(kd - kShift, kNegLn2OverN, x);
    const double t = fma(r * 0.5, r, r);
    const double S = exp_table_scale(kd, tab);
    return fma(S, t, S);
}

ALP_HD double exp_dec(double x, const double* tab = nullptr) {
    const double kShift = 6755399441055744.0;
    const double kd = fma(x, ALP_EXPC(0), kShift);
    const double r = fma(kd - kShift, kNegLn2OverN, x);
    return exp_table_finish(kd, r, tab);
}

// 10^x with the same table and polynomial: x = (N n + j) log10(2)/N + t, 10^x = 2^n 2^(j/N) exp(t ln10).
// Requires |x| <= 300.  ~1 ulp; used for the log-uniform energy draw of the bremsstrahlung channel (fluxes.py:326-327).
ALP_HD double exp10_fast_core(double x, const double* tab = nullptr) {
    const double kShift = 6755399441055744.0;
    const double kd = fma(x, ALP_EXP10C(0), kShift);                // N x log2(10)
    const double fn = kd - kShift;
    double t = fma(fn, ALP_EXP10C(1), x);                           // log10(2)/N, high part (32 trailing zero bits)
    t = fma(fn, ALP_EXP10C(2), t);                                  // low part
    return exp_table_finish(kd, t * ALP_EXP10C(3), tab);            // |r| <= ln2/(2N)
}

ALP_HD double exp_fast(double x) {
    if (!(fabs(x) <= 708.0)) return exp(x);
    return exp_fast_core(x);
}

// ---- np.heaviside ---------------------------------------------------------------------------------------------
ALP_HD double heaviside(double x, double h0) {
    if (x != x) return x;            // NaN propagates
    return x > 0.0 ? 1.0 : (x < 0.0 ? 0.0 : h0);
}

// ---- np.interp (numpy/_core/src/multiarray/compiled_base.c arr_interp) ------------------------------------------
// xp strictly increasing, n >= 1.
ALP_HD double interp(double x, const double* xp, const double* fp, int n, double left, double right) {
    if (x != x) return x;
    if (x < xp[0]) return left;
    if (x > xp[n - 1]) return right;
    int lo = 0, hi = n - 1;          // invariant xp[lo] <= x <= xp[hi]
    if (x == xp[n - 1]) return fp[n - 1];
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= xp[mid]) lo = mid; else hi = mid;
    }
    if (xp[lo] == x) return fp[lo];
    const double slope = (fp[lo + 1] - fp[lo]) / (xp[lo + 1] - xp[lo]);
    double r = slope * (x - xp[lo]) + fp[lo];
    if (r != r) {
        r = slope * (x - xp[lo + 1]) + fp[lo + 1];
        if (r != r && fp[lo] == fp[lo + 1]) r = fp[lo];
    }
    return r;
}

// photon_xs.py:48-49  AbsCrossSection.sigma_mev: 10**interp(log10 E; log10 E_i, log10(dim*sigma_i), 0, 0)/MEV2_CM2.
// (Out-of-range energies give 10**0/MEV2_CM2 -- the reference's behaviour)
ALP_HD double abs_sigma_mev(double e, const double* log10_e, const double* log10_xs, int n) {
    const double y = interp(log10(e), log10_e, log10_xs, n, 0.0, 0.0);
    return pow(10.0, y) / kMeV2Cm2;
}

// ---- prod_xs.py:99-104  primakoff_sigma -------------------------------------------------------------------------
// pref = (g*z)**2/(2*137) and r02 = r0**2 are evaluated by the host (Python-float arithmetic, like the reference).
ALP_HD double primakoff_sigma(double eg, double ma, double pref, double r02) {
    const double eta2 = r02 * (eg * eg);
    return heaviside(eg - ma, 0.0) * pref * (((2 * eta2 + 1) / (4 * eta2)) * log(1 + 4 * eta2) - 1);
}

// ---- det_xs.py:37-42  iprimakoff_sigma ----------------------------------------------------------------------------
ALP_HD double iprimakoff_sigma(double ea, double ma, double ma2, double pref, double r02) {
    const double eta2 = r02 * (ea * ea - ma2);
    return heaviside(ea - ma, 0.0) * pref * (((2 * eta2 + 1) / (4 * eta2)) * log(1 + 4 * eta2) - 1);
}

// ---- det_xs.py:89-98  icompton_sigma --------------------------------------------------------------------------------
struct IComptonConst {
    double ma, ma2, ma4;    // ma, ma**2, ma**4
    double ethr;            // ma*sqrt(2*M_E**2 + ma**2)/(2*M_E)
    double zfac;            // z*ALPHA*power(g/M_E,2)  (left-to-right product after the two heavisides)
    double c2;              // ma**4 + 2*power(ma*M_E,2)
};

ALP_HD double icompton_sigma(double ea, const IComptonConst& c) {
    const double me = kME, me2 = kME * kME;
    const double y = (2 * me) * ea + c.ma2;
    double d = ea * ea - c.ma2;
    if (d < 0.0) d = 0.0;                                        // np.clip(..., a_min=0)
    const double pa = sqrt(d);
    const double pref = ((heaviside(ea - c.ma, 0.0) * heaviside(ea - c.ethr, 0.0)) * c.zfac) / (8 * pa);
    const double m2y = me2 + y;
    const double tme = (2 * me) * ea;
    const double tmp = (2 * me) * pa;
    const double t1 = (((2 * me2) * (me + ea)) * y) / (m2y * m2y);
    const double t2 = ((4 * me) * (c.c2 - tme * tme)) / (y * m2y);
    const double t3 = (log(((me + ea) + pa) / ((me + ea) - pa)) * (tmp * tmp + c.ma4)) / (pa * y);
    double r = pref * ((t1 + t2) + t3);
    // np.nan_to_num: nan -> 0, +-inf -> +-DBL_MAX
    if (r != r) return 0.0;
    if (isinf(r)) return r > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return r;
}

// Fast mode of the two detection cross sections (<= 1e-9 relative from the reference expressions above): the seven
// divisions of icompton_sigma become four reciprocals (hardware seed + one correction,
// no special-case branches); samples outside the straight-line domain (E_a <= m_a, below threshold) take the reference
// expression, which also reproduces its nan_to_num corner cases.
ALP_HD double icompton_sigma_fast(double ea, const IComptonConst& c) {
    const double d = ea * ea - c.ma2;
    if (!(d > 1e-290) || !(ea < 1e140) || !(ea > 0.0)) return icompton_sigma(ea, c);       // (rare: keeps warps converged)
    const double me = kME, me2 = kME * kME;
    const double y = (2 * me) * ea + c.ma2;
    // p_a keeps its correctly rounded square root: (m_e + E_a) - p_a cancels by a factor ~E_a/m_e, so an ulp of p_a would be
    // amplified beyond 1e-9 above ~1 TeV; with the same p_a the difference is the reference's, bit for bit
    const double pa = sqrt(d);
    const double r_pa = rcp_fast(pa);
    const double m2y = me2 + y;
    const double r_y = rcp_fast(y), r_m = rcp_fast(m2y);
    const double A = me + ea;
    const double tme = (2 * me) * ea, tmp = (2 * me) * pa;
    const double t1 = ((((2 * me2) * A) * y) * r_m) * r_m;
    const double t2 = (((4 * me) * (c.c2 - tme * tme)) * r_y) * r_m;
    const double t3 = ((log((A + pa) * rcp_fast(A - pa)) * (tmp * tmp + c.ma4)) * r_pa) * r_y;
    // below threshold the reference's prefactor is heaviside(.) * zfac / (8 p_a) = 0 and nan_to_num maps 0 * (anything) to 0
    return ea > c.ethr ? ((c.zfac * 0.125) * r_pa) * ((t1 + t2) + t3) : 0.0;
}

ALP_HD double iprimakoff_sigma_fast(double ea, double ma, double ma2, double pref, double r02) {
    const double eta2 = r02 * (ea * ea - ma2);
    // (eta2 << 1, i.e. within ~1e-8 of threshold: the bracket cancels like 1/eta2 and amplifies the last bit of the
    // quotient -- keep the reference's own division there)
    if (!(ea > ma) || !(eta2 > 1e-2) || !(eta2 < 1e140)) return iprimakoff_sigma(ea, ma, ma2, pref, r02);
    return pref * (((2 * eta2 + 1) * (0.25 * rcp_fast(eta2))) * log(1 + 4 * eta2) - 1);
}

// ---- Compton production: per-bin constants (prod_xs.py:155-169, fluxes.py:210-224) -----------------------------------
// Layout of one row of the per-bin parameter table (doubles).
enum ComptonBin { CB_EG = 0, CB_SPAN, CB_C0, CB_XMIN, CB_XMAX, CB_P, CB_K, CB_S, CB_SIGABS, CB_PHI, CB_Q, CB_INVEG, CB_NP };

struct ComptonGlobals {
    double ma, ma2;      // ma, ma**2 (Python float pow on the host)
    double aa;           // g**2/4/pi
    double z;            // target Z
    double n_samples;    // as double (divisor)
};

ALP_HD void compton_bin_params(double eg, double phi, double sigabs, const ComptonGlobals& g, double* out) {
    const double me2 = kME * kME;
    const double s = (2 * kME) * eg + me2;                        // :158
    const double A = s - me2;
    const double B = A + g.ma2;
    const double D = sqrt(B * B - (4 * s) * g.ma2);
    const double den = (2 * s) * A;
    out[CB_EG] = eg;
    out[CB_SPAN] = eg - g.ma;
    out[CB_C0] = g.ma2 / ((2 * eg) * kME);                        // :159
    out[CB_XMIN] = (A * B - A * D) / den;                         // :161-162
    out[CB_XMAX] = (A * B + A * D) / den;                         // :163-164
    // z * thresh * (1/eg) * pi * a * aa / (s - M_E**2), thresh factored out (0 or 1)    :167
    out[CB_P] = (((((g.z * heaviside(eg - g.ma, 0.0)) * (1 / eg)) * kPi) * (1.0 / 137.0)) * g.aa) / A;
    out[CB_K] = (-2 * g.ma2) / (A * A);
    out[CB_S] = s;
    out[CB_SIGABS] = sigabs;
    out[CB_PHI] = phi;
    // fast mode: every per-row factor of the weight folded into one constant
    out[CB_Q] = phi * (((out[CB_SPAN] * out[CB_P]) / g.n_samples) / sigabs);
    out[CB_INVEG] = 1 / eg;
}

// One Compton sample: returns E_a, writes the production weight (fluxes.py:218-224).
ALP_HD double compton_sample(double u, const double* b, const ComptonGlobals& g, double& flux) {
    const double me2 = kME * kME;
    const double eg = b[CB_EG];
    const double ea = g.ma + b[CB_SPAN] * u;                      // np.random.uniform(ma, eg): lo + (hi-lo)*u
    const double x = (b[CB_C0] - ea / eg) + 1;                    // prod_xs.py:159
    const double th = heaviside(x - b[CB_XMIN], 0.0) * heaviside(b[CB_XMAX] - x, 0.0);   // :166
    const double omx = 1 - x;
    const double inner = b[CB_K] * ((b[CB_S] - me2 / omx) - g.ma2 / x) + x;              // :167-168
    const double dsig = (b[CB_P] * th) * ((x / omx) * inner);
    const double mc_xs = (b[CB_SPAN] * dsig) / g.n_samples;       // fluxes.py:219
    flux = b[CB_PHI] * (mc_xs / b[CB_SIGABS]);                    // :220, :224
    return ea;
}

// Fast mode (same quantity, <= 1e-9 relative from the reference): x keeps its exact division (it is the
// cancellation-sensitive variable), the three remaining divisions share one reciprocal 1/(x(1-x)), and the per-row
// factors are pre-folded into CB_Q.
ALP_HD double compton_sample_fast(double u, const double* b, const ComptonGlobals& g, double& flux) {
    const double me2 = kME * kME;
    const double ea = g.ma + b[CB_SPAN] * u;
    // ea/eg by reciprocal + one fused correction step: the correctly rounded quotient in all but rare last-bit cases
    // (x is the cancellation-sensitive variable, so its quotient is kept to that accuracy)
    const double q0 = ea * b[CB_INVEG];
    const double q = fma(fma(-q0, b[CB_EG], ea), b[CB_INVEG], q0);
    const double x = (b[CB_C0] - q) + 1;
    const bool in = (x > b[CB_XMIN]) && (x < b[CB_XMAX]);
    const double omx = 1 - x;
    const double t = rcp_fast1(x * omx);
    const double r_omx = x * t, r_x = omx * t;                    // 1/(1-x), 1/x
    const double inner = fma(b[CB_K], fma(-g.ma2, r_x, fma(-me2, r_omx, b[CB_S])), x);
    // outside the window the reference's prefactor is an exact 0 (its heavisides) that multiplies the bracket: select the
    // per-row factor (Q or 0*Q, both row constants -- hoisted out of the sample loop where the row is fixed) instead of
    // multiplying the result by 0; NaN/inf brackets still propagate as 0*inf = NaN
    const double qrow = in ? b[CB_Q] : 0.0 * b[CB_Q];
    flux = qrow * ((x * r_omx) * inner);
    return ea;
}

// ---- fluxes.py:43-65 + 159-162  propagate ----------------------------------------------------------------------------
struct PropConst {
    double ma, ma2;           // ma, ma**2
    double width;             // decay width; <= 0 -> infinite lifetime (fluxes.py:51)
    double rescale;           // rescale_factor
    double neg_dist_mbm;      // -det_dist / METER_BY_MEV        (host-evaluated, Python floats)
    double neg_len_mbm;       // -det_length / METER_BY_MEV
    double geom;              // det_area/(4*pi*det_dist**2) or unused
    float  geom32;            // float32(geom)  (NEP 50: Python-float factor multiplies a float32 array in float32)
    int    apply_geom;        // is_isotropic
    int    geom_f64;          // factor was an np.float64 -> multiply in f64 then cast
    int    use_tof;           // time_of_flight_cut is not None
    double tof_light;         // det_dist / (1e-2*C_LIGHT)
    double det_dist;          // metres
    double timing_window;     // seconds
    int    fast;              // 1: fast mode, one rsqrt instead of sqrt + 7 divisions (<= 1e-9 relative); 2: + FP32 transcendentals
    double k1, k2;            // fast mode: neg_dist_mbm*ma*width, neg_len_mbm*ma*width (0 for width <= 0)
    double t_safe;            // fast mode: 708/max(|k1|,|k2|) -- largest 1/p_a for which exp_fast_core applies
    int    short_lived;       // fast mode: exponents reach -708 at p_a >= 1 MeV (see prop_derive / propagate_try_fast)
    double t_fast;            // min(t_safe, 1e140) when the branch-free path applies (fast == 1, no time-of-flight cut), else -1
};

// Derived fields of PropConst (after the plain ones are filled in).
ALP_HD void prop_derive(PropConst& c) {
    const double m = fmax(fabs(c.k1), fabs(c.k2));
    c.t_safe = m > 0.0 ? 708.0 / m : INFINITY;
    c.t_fast = (c.fast == 1 && !c.use_tof) ? fmin(c.t_safe, 1e140) : -1.0;
    // short-lived regime: exponents fall below -708 for p_a < 1/t_safe; once that is 1 MeV or more a sizeable share of the
    // samples is affected -> the launchers use the SHORT_LIVED kernels
    c.short_lived = (c.fast == 1 && !c.use_tof && c.t_safe < 1.0) ? 1 : 0;
}

// Returns the FP64 weights before the float32 cast (decay, scatter); the casts + acceptance are prop_cast().
ALP_HD void propagate_sample(double ea, double wgt, const PropConst& c, double& d64, double& s64) {
    const double pa = sqrt(ea * ea - c.ma2);                      // :48
    const double va = pa / ea;                                    // :49
    const double boost = ea / c.ma;                               // :50
    const double tau = c.width > 0.0 ? boost / c.width : INFINITY;   // :51
    if (c.use_tof) {                                              // :54-58
        const double tof = c.det_dist / ((va * 1e-2) * kCLight);
        const double dt = fabs(c.tof_light - tof);
        wgt = wgt * ((dt < c.timing_window) ? 1.0 : 0.0);
    }
    const double surv = exp((c.neg_dist_mbm / va) / tau);         // :61
    const double dec = 1 - exp((c.neg_len_mbm / va) / tau);       // :62
    s64 = (c.rescale * wgt) * surv;                               // :65
    d64 = s64 * dec;                                              // :64
}

// Fast mode: L/(hbar c v tau) = L ma Gamma / (hbar c p_a), so both exponents are a constant times 1/p_a.
// Branch-free attempt: valid (returns true) for the plain FP64 fast mode when p_a^2 is a positive normal number and both
// exponents lie in [-708, 0]; otherwise the outputs are garbage and propagate_sample_fast() redoes the sample.
// (d <= 1e-300, negative or NaN gives t >= 1e150, inf or NaN: t_fast <= 1e140 rejects them with the one comparison.)
// SHORT_LIVED: variant for decay lengths far below the baseline, where a sizeable fraction of the samples has exponents
// below -708.  Each exponential is then handled on its own: below -746 it is exactly 0 (survival 0 / decay probability 1),
// so only the narrow band (-746, -708) of denormal results still goes to the library path.  The launchers pick the
// variant from the host-known exponent scale (PropConst::short_lived); the long-lived regime pays nothing for it.
template <bool SHORT_LIVED = false, bool PLAIN = false>      // PLAIN: the launcher guarantees c.fast == 1 (no FP32 mode)
ALP_HD bool propagate_try_fast(double ea, double wgt, const PropConst& c, double& d64, double& s64,
                               const double* exp_tab = nullptr) {
    const double d = ea * ea - c.ma2;          // (NOT fused: near threshold the reference's rounding of ea*ea dominates p_a)
    if (!PLAIN && c.fast == 2) {
        // optional FP32 mode (north star: 1e-5 relative): only the transcendental part is single precision -- rsqrt and
        // the two exponentials (~1e-7 relative each); the decay probability uses expm1f, which unlike the reference's
        // 1 - exp(-x) does not cancel for long-lived ALPs.  Products stay FP64 so the dynamic range is unchanged.
        const float tf = rsqrtf((float)d);
        const float surv_f = expf((float)c.k1 * tf);
        const float dec_f = -expm1f((float)c.k2 * tf);
        s64 = (c.rescale * wgt) * (double)surv_f;
        d64 = s64 * (double)dec_f;
        return (d > 0.0) & (c.use_tof == 0);
    }
    const double t = rsqrt_fast1(d);
    if (!SHORT_LIVED) {
        const double surv = exp_surv(c.k1 * t, exp_tab);
        const double dec = 1 - exp_dec(c.k2 * t, exp_tab);
        s64 = (c.rescale * wgt) * surv;
        d64 = s64 * dec;
        return t <= c.t_fast;
    }
    const double x1 = c.k1 * t, x2 = c.k2 * t;                     // both <= 0 (or NaN)
    const bool z1 = x1 < -746.0, z2 = x2 < -746.0;
    const double e1 = exp_surv(z1 ? -1.0 : x1, exp_tab), e2 = exp_dec(z2 ? -1.0 : x2, exp_tab);
    const double surv = z1 ? 0.0 : e1;
    const double dec = 1 - (z2 ? 0.0 : e2);
    s64 = (c.rescale * wgt) * surv;
    d64 = s64 * dec;
    return (z1 | (x1 >= -708.0)) & (z2 | (x2 >= -708.0)) & (c.t_fast >= 0.0);      // (NaN exponents fail both tests)
}

ALP_HD void propagate_sample_fast(double ea, double wgt, const PropConst& c, double& d64, double& s64) {
    // (the same variant as the pair kernels of this launch, so every kernel gives the same bits for a sample)
    if (c.short_lived ? propagate_try_fast<true>(ea, wgt, c, d64, s64) : propagate_try_fast<false>(ea, wgt, c, d64, s64)) return;
    const double d = ea * ea - c.ma2;
    if (!(d > 0.0) || c.use_tof) {            // E_a <= m_a (v = 0 or NaN) and the time-of-flight cut: reference path
        propagate_sample(ea, wgt, c, d64, s64);
        return;
    }
    const double t = rsqrt(d);                // exponents below -708 (or a subnormal p_a^2): library functions
    const double surv = exp(c.k1 * t);
    const double dec = 1 - exp(c.k2 * t);
    s64 = (c.rescale * wgt) * surv;
    d64 = s64 * dec;
}

ALP_HD float prop_cast(double w64, const PropConst& c) {
    const float w = (float)w64;                                   // np.asarray(..., dtype=np.float32)
    // fluxes.py:159-162.  geom32 is 1.0f when is_isotropic is False (an exact no-op), so the common case is one FMUL.
    if (c.geom_f64) return (float)((double)w * c.geom);        // (set only together with apply_geom)
    return w * c.geom32;
}

// Fast mode: always the float32 multiply (an np.float64 acceptance factor would be applied in float64 by the reference:
// <= 1 float32 ulp apart, the tolerance of the float32 attributes).
ALP_HD float prop_cast_fast(double w64, const PropConst& c) { return (float)w64 * c.geom32; }

// ---- generators.py:88-92  decays: days*S_PER_DAY*w32*heaviside(E-thr,1) ---------------------------------------------------
struct EventConst {
    double days_sec;      // days_exposure*S_PER_DAY
    float  days_sec32;    // float32 of it (Python-scalar * float32 array stays float32 under NEP 50)
    int    days_f64;      // exposure was an np.float64 -> product in f64
    double threshold;
};

ALP_HD double decay_event_weight(double ea, float w32, const EventConst& c) {
    const double base = c.days_f64 ? c.days_sec * (double)w32 : (double)(c.days_sec32 * w32);
    const double d = ea - c.threshold;
    return base * (d >= 0.0 ? 1.0 : (d < 0.0 ? 0.0 : d));        // heaviside(d, 1.0); NaN propagates
}

// The same inside the fused chain, where E_a and the weight come from the same sample (a NaN E_a implies a NaN weight, so
// the NaN branch of heaviside is covered by the product) and E_a - threshold >= 0 <=> E_a >= threshold (IEEE subtraction
// is zero only for equal operands).
ALP_HD double decay_event_weight_coupled(double ea, float w32, const EventConst& c) {
    if (c.days_f64) {
        const double base = c.days_sec * (double)w32;
        return ea >= c.threshold ? base : base * 0.0;
    }
    // Python-scalar exposure: the product and the threshold factor live in float32 (NEP 50); the widening is exact, so the
    // select can be done before it -- one FMUL instead of a DMUL for the 0 * base that keeps NaN / inf weights NaN
    const float b32 = c.days_sec32 * w32;
    return (double)(ea >= c.threshold ? b32 : b32 * 0.0f);
}

// PLAIN variants for kernels whose launcher has excluded the rare modes (np.float64 exposure, FP32 transcendentals): the mode
// tests disappear from the sample loop.
ALP_HD double decay_event_weight_plain(double ea, float w32, const EventConst& c) {
    const float b32 = c.days_sec32 * w32;
    return (double)(ea >= c.threshold ? b32 : b32 * 0.0f);
}

// ---- 2-body decay a -> gamma gamma of a forward parent (generators.py:116-126, cross_section_mc.py:510-544,669-685) ----------
enum ParentRow { PR_ECM = 0, PR_PCM, PR_GAM, PR_M03, PR_M33, PR_W, PR_NP };

// Per-parent constants.  ma2 = ma**2.  Returns false when the boost is the identity (beta == 0).
ALP_HD void decay_parent_params(double ea, double ma2, double w, double* out) {
    const double pa = sqrt(ea * ea - ma2);                        // generators.py:118
    const double pz = pa;                                         // pa*cos(0.0)
    const double mp = sqrt(ea * ea - pz * pz);                    // LorentzVector.mass(): sqrt(dot(pmu**2, mt))
    const double mp2 = mp * mp;
    const double pcm = sqrt((mp2 - 0.0) * (mp2 - 0.0)) / (2 * mp);   // cross_section_mc.py:524 (m1 = m2 = 0)
    const double ecm = sqrt(pcm * pcm + 0.0);                     // :525
    const double vz = -(pz / ea);                                 // :532
    const double beta = sqrt(vz * vz);                            // Vector3.mag
    double gam = 1.0, m03 = 0.0, m33 = 1.0;
    if (beta != 0.0) {                                            // lorentz_boost :678-685
        const double nz = vz / beta;
        gam = 1 / sqrt(1 - beta * beta);
        m03 = ((-gam) * beta) * nz;
        m33 = 1 + ((gam - 1) * nz) * nz;
    }
    out[PR_ECM] = ecm; out[PR_PCM] = pcm; out[PR_GAM] = gam; out[PR_M03] = m03; out[PR_M33] = m33; out[PR_W] = w;
}

// One decay: u1 -> phi, u2 -> theta.  p1/p2 = (E, px, py, pz) lab frame.
ALP_HD void decay_event(double u1, double u2, const double* p, double* p1, double* p2) {
    const double phi = kTwoPi * u1;                               // :529
    const double th = acos(1 - 2 * u2);                           // :530
    double sphi, cphi, sth, cth;
#if defined(__CUDA_ARCH__)
    sincos(phi, &sphi, &cphi);
    sincos(th, &sth, &cth);
#else
    sphi = sin(phi); cphi = cos(phi); sth = sin(th); cth = cos(th);
#endif
    const double pcm = p[PR_PCM], ecm = p[PR_ECM], gam = p[PR_GAM], m03 = p[PR_M03], m33 = p[PR_M33];
    const double px = (pcm * cphi) * sth;                         // :535-537
    const double py = (pcm * sphi) * sth;
    const double pz = pcm * cth;
    p1[0] = gam * ecm + m03 * pz;                                 // mat @ pmu, rows 0 and 3
    p1[1] = px; p1[2] = py;
    p1[3] = m03 * ecm + m33 * pz;
    p2[0] = gam * ecm + m03 * (-pz);
    p2[1] = -px; p2[2] = -py;
    p2[3] = m03 * ecm + m33 * (-pz);
}

// ---- general 2-body decay: parent with arbitrary direction, daughter masses m1, m2 -------------------------------------------
// (Decay2Body.__init__/decay cross_section_mc.py:496-544 with the dense 4x4 lorentz_boost :669-685; used by parents
// with a polar angle, generators.py:117-121, and by FluxNeutralMeson2BodyDecay fluxes.py:1026-1031)
enum ParentGRow { PG_MAT = 0, PG_E1 = 16, PG_E2, PG_PCM, PG_W, PG_NP };

ALP_HD void decay_parent_params_general(double e, double px, double py, double pz, double m1, double m2, double w,
                                        double* out) {
    const double mp = sqrt(((e * e - px * px) - py * py) - pz * pz);              // LorentzVector.mass() :113-114
    const double mp2 = mp * mp;
    const double dm = m2 - m1, sm = m2 + m1;
    const double pcm = sqrt((mp2 - dm * dm) * (mp2 - sm * sm)) / (2 * mp);         // :524
    out[PG_E1] = sqrt(pcm * pcm + m1 * m1);                                        // :525
    out[PG_E2] = sqrt(pcm * pcm + m2 * m2);                                        // :526
    out[PG_PCM] = pcm;
    out[PG_W] = w;
    const double vx = -(px / e), vy = -(py / e), vz = -(pz / e);                   // :532  v_in = -velocity
    const double beta = sqrt((vx * vx + vy * vy) + vz * vz);                       // Vector3.mag :59-60
    double* m = out + PG_MAT;
    if (beta == 0.0) {                                                             // :678-679 momentum returned unchanged
        for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0;
        return;
    }
    double n0 = 0.0, n1 = 0.0, n2 = 0.0;
    if (!(beta < 1e-40)) { n0 = vx / beta; n1 = vy / beta; n2 = vz / beta; }       // unit_vec :49-53
    const double g = 1 / sqrt(1 - beta * beta);                                    // :680
    const double gb = (-g) * beta, gm1 = g - 1;
    m[0] = g;        m[1] = gb * n0;              m[2] = gb * n1;              m[3] = gb * n2;
    m[4] = gb * n0;  m[5] = 1 + (gm1 * n0) * n0;  m[6] = (gm1 * n0) * n1;      m[7] = (gm1 * n0) * n2;
    m[8] = gb * n1;  m[9] = (gm1 * n1) * n0;      m[10] = 1 + (gm1 * n1) * n1; m[11] = (gm1 * n1) * n2;
    m[12] = gb * n2; m[13] = (gm1 * n2) * n0;     m[14] = (gm1 * n2) * n1;     m[15] = 1 + (gm1 * n2) * n2;
}

ALP_HD void boost_apply(const double* m, double p0, double p1, double p2, double p3, double* out) {
#pragma unroll
    for (int r = 0; r < 4; ++r)                                                    // mat @ pmu, row by row
        out[r] = ((m[4 * r] * p0 + m[4 * r + 1] * p1) + m[4 * r + 2] * p2) + m[4 * r + 3] * p3;
}

// One decay; q1/q2 = lab 4-vectors (E, px, py, pz); returns theta of daughter 1 (LorentzVector.theta :131-132)
ALP_HD double decay_event_general(double u1, double u2, const double* row, double* q1, double* q2) {
    const double phi = kTwoPi * u1;
    const double th = acos(1 - 2 * u2);
    const double pcm = row[PG_PCM];
    const double px = (pcm * cos(phi)) * sin(th), py = (pcm * sin(phi)) * sin(th), pz = pcm * cos(th);
    boost_apply(row + PG_MAT, row[PG_E1], px, py, pz, q1);
    boost_apply(row + PG_MAT, row[PG_E2], -px, -py, -pz, q2);
    const double pm = sqrt((q1[1] * q1[1] + q1[2] * q1[2]) + q1[3] * q1[3]);
    return acos(q1[3] / pm);
}

// Fast mode: cos(theta) = 1-2u and sin(theta) = 2 sqrt(u(1-u)) directly (no acos -> sincos round trip) and
// sincospi(2 u1) for the azimuth.  |dp|/E_parent vs the reference ~1e-15.
ALP_HD void decay_event_fast(double u1, double u2, const double* p, double* p1, double* p2) {
    double sphi, cphi;
#if defined(__CUDA_ARCH__)
    sincospi(2 * u1, &sphi, &cphi);
#else
    sphi = sin(kTwoPi * u1); cphi = cos(kTwoPi * u1);
#endif
    const double cth = 1 - 2 * u2;
    const double sth = 2 * sqrt(u2 * (1 - u2));
    const double pcm = p[PR_PCM], ecm = p[PR_ECM], gam = p[PR_GAM], m03 = p[PR_M03], m33 = p[PR_M33];
    const double px = (pcm * cphi) * sth;
    const double py = (pcm * sphi) * sth;
    const double pz = pcm * cth;
    const double ge = gam * ecm, me = m03 * ecm, mp = m03 * pz, m3p = m33 * pz;
    p1[0] = ge + mp; p1[1] = px; p1[2] = py; p1[3] = me + m3p;
    p2[0] = ge - mp; p2[1] = -px; p2[2] = -py; p2[3] = me - m3p;
}

}  // namespace alp
