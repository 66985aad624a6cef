// Per-sample arithmetic of the charged-meson 3-body decay fluxes  M -> l nu X  (X = dark scalar / pseudoscalar / vector):
// matrix elements (matrix_element.py:50-68, 419-518), the Dalitz integral dGamma/dE_a (fluxes.py:639-661, 841-863) and
// the sample kinematics of FluxChargedMeson3BodyIsotropic (fluxes.py:877-902) and FluxChargedMeson3BodyDecay
// (fluxes.py:669-725).  Same conventions as alp_math.cuh: host+device, reference op order, no FMA contraction.
//
// The reference integrates the matrix element with scipy.integrate.quad (QUADPACK QAGS, epsabs = epsrel = 1.49e-8).  With
// the couplings this path is used at, the integrals are << epsabs, so QAGS returns after its first 21-point
// Gauss-Kronrod evaluation of the whole interval; qags21() reproduces exactly that evaluation and QAGS's acceptance
// test, and continues with QAGS's own bisection rule (largest error estimate first) where the test fails.  QAGS's
// epsilon-algorithm extrapolation is not reproduced: it only comes into play from the third bisection level on, where
// the two agree within the reference's own tolerance (1.49e-8) instead of 1e-9.
#pragma once
#include "alp_math.cuh"

namespace alp {

constexpr double kMMu = 105.658369;        // constants.py:34
constexpr double kGF = 1.16638e-11;        // constants.py:73
constexpr double kCabibbo = 0.9743;        // constants.py:74

enum Meson3Model : int {
    kM3ScalarIB1 = 0, kM3PseudoscalarIB1, kM3VectorIB1, kM3ScalarIB2, kM3VectorIB2, kM3VectorIB9, kM3SD, kM3VectorContact,
    kM3None
};

struct Meson3Const {
    double mm;              // parent meson mass
    double ma;              // dark boson mass (m3)
    double fM, ckm;         // decay constant, CKM element
    double total_width;
    double c0;              // contact-model parameter
    double coupling;
    double a, b, d;         // vector_ib9 free parameters
    double pref;            // (2*mm)/(32*power(2*pi*mm, 3))      (host-evaluated, Python floats)
    double n_samples;       // as double (divisor)
    int model;
};

// MatrixElementDecay3.get_sp_from_dalitz (matrix_element.py:60-68), m2 = 0
struct SP6 { double sp01, sp02, sp03, sp12, sp13, sp23; };

ALP_HD SP6 m3_scalar_products(double M, double m1, double m3, double m122, double m232) {
    SP6 s;
    s.sp03 = ((M * M + m3 * m3) - m122) / 2;
    s.sp01 = ((M * M + m1 * m1) - m232) / 2;
    s.sp02 = (((m122 + m232) - m1 * m1) - m3 * m3) / 2;
    s.sp13 = (((M * M + 0.0 * 0.0) - m122) - m232) / 2;
    s.sp23 = ((m232 - 0.0 * 0.0) - m3 * m3) / 2;
    s.sp12 = ((m122 - m1 * m1) - 0.0 * 0.0) / 2;
    return s;
}

ALP_HD double cube(double x) { return (x * x) * x; }

// The part of the *_ib1 matrix elements that depends on m122 only (matrix_element.py:446, 451): the threshold factor
// heaviside(e3star - m3, 0).  Evaluated once per Dalitz integral instead of once per node (same value, bit for bit).
ALP_HD double m2_threshold_factor(double m122, const Meson3Const& K) {
    if (K.model != kM3ScalarIB1 && K.model != kM3PseudoscalarIB1) return 1.0;
    const double e3star = ((K.mm * K.mm - m122) - K.ma * K.ma) / (2 * sqrt(m122));
    return heaviside(e3star - K.ma, 0.0);
}

// M2Meson3BodyDecay.__call__ (matrix_element.py:443-518).  m1 = lepton mass; hv = m2_threshold_factor(m122, K).
ALP_HD double m2_meson3body(double m122, double m232, double m1, const Meson3Const& K, double hv) {
    const double M = K.mm, m3 = K.ma, g = K.coupling;
    const double M2 = M * M, m12 = m1 * m1, m32 = m3 * m3;
    const SP6 sp = m3_scalar_products(M, m1, m3, m122, m232);
    // (lp, pq, kp, lq, kl, kq) = (sp01, sp02, sp03, sp12, sp13, sp23)       :444
    const double lp = sp.sp01, pq = sp.sp02, kp = sp.sp03, lq = sp.sp12, kl = sp.sp13, kq = sp.sp23;
    switch (K.model) {
    case kM3ScalarIB1:
    case kM3PseudoscalarIB1: {                                                        // :445-464
        const double ev = (((m122 + m232) - m12) - m32) / (2 * M);
        const double emu = ((M2 - m232) + m12) / (2 * M);
        const double q2 = M2 - (2 * M) * ev;
        const double r = (((g * kGF) * K.fM) * K.ckm) / (q2 - m12);
        const double prefactor = hv * (r * r);
        const double m1M = m1 * M;
        const double head = (((2 * M) * emu) * q2) * (q2 - m12) - (q2 * q2 - m1M * m1M) * ((q2 + m12) - m32);
        const double tail = ((2 * q2) * m12) * (M2 - q2);
        return K.model == kM3ScalarIB1 ? prefactor * (head + tail) : prefactor * (head - tail);
    }
    case kM3VectorIB1: {                                                              // :465-472
        const double q2 = M2 - ((2 * M) * (((m122 + m232) - m12) - m32)) / (2 * M);
        const double r = (((kGF * K.fM) * K.ckm) / (q2 - m12)) / m3;
        const double prefactor = 8 * (r * r);
        // second unpacking under different names (:468): kl, kq, kp, lq, lp, pq = sp01, sp02, sp03, sp12, sp13, sp23
        const double kl2 = sp.sp01, kq2 = sp.sp02, kp2 = sp.sp03, lq2 = sp.sp12, lp2 = sp.sp13, pq2 = sp.sp23;
        const double cr = g, cl = g;
        const double crMm = (cr * M) * m1, clq = cl * q2;
        return -prefactor * ((crMm * crMm - clq * clq) * (lq2 * m32 + (2 * lp2) * pq2)
                             - (((2 * cr) * m12) * kq2) * (((cr * m32) * kl2 + ((2 * cr) * kp2) * lp2) - ((3 * cl) * q2) * m32));
    }
    case kM3ScalarIB2:                                                                // :473-474
        return 0.0;
    case kM3VectorIB2: {                                                              // :475-478
        const double t = ((g * kGF) * K.fM) * m1;
        const double prefactor = 2 * (t * t);
        const double u = (M2 - m122) + m32;
        const double dn = M2 - m122;
        return (-prefactor * ((m12 - m122) * (u * u - (4 * M2) * m32))) / (m32 * (dn * dn));
    }
    case kM3VectorIB9: {                                                              // :479-482, 500-518 (MIB9, F = 1)
        const double a = K.a, b = K.b, d = K.d, mV2 = m32, ml2 = m12, F2 = 1.0;
        const double A = (M2 - m122) - m232;              // M**2 + -1*m122 + -1*m232
        const double B = m232 - mV2;                      // m232 + -1*mV**2
        const double Cn = (-M2 + m122) + m232;            // -1*M**2 + m122 + m232
        const double Dm = m122 - ml2;                     // m122 + -1*ml**2
        const double Sp = m122 + ml2;
        double res = ((b * b) * cube(A)) * B;
        res = res + (((2 * b) * d) * ml2) * cube(B);
        res = res + (((4 * b) * (a * m122 + (a + (2 * d) * m122) * ml2)) * mV2) * (-m232 + mV2);
        res = res + ((-4 * Dm) * mV2) * (((-(a * a) + ((2 * a) * d) * ml2) + ((d * d) * m122) * ml2)
                                         + ((-(b * b)) * m122 + F2 * Sp) * mV2);
        res = res + (B * B) * ((((4 * a) * d) * ml2 + ((d * d) * ml2) * Dm) + ((b * b) * (-m122 + ml2) + (2 * F2) * Sp) * mV2);
        res = res + (2 * (Cn * Cn)) * (((b * ((2 * a) + d * ml2)) * B + (b * b) * (B * B))
                                       + (0.5 * Dm) * ((d * d) * ml2 + (-((b * b) + -2 * F2)) * mV2));
        res = res + A * (((((4 * a) * b) * (-m122 + ml2)) * mV2 + ((4 * b) * (a + d * ml2)) * (B * B)) + (b * b) * cube(B)
                         + (2 * B) * ((((2 * a) * a + ((2 * a) * d) * ml2) + ((d * d) * ml2) * Dm) + ((b * b) * (-3 * m122 + ml2)) * mV2));
        res = (0.5 * (1.0 / mV2)) * res;
        return (((g * g) * res) * (kGF * kGF)) * (K.fM * K.fM);
    }
    case kM3SD: {                                                                     // :483-490
        const double t = ((g * kGF) * kCabibbo) / sqrt(2.0) / M;
        const double prefactor = 8 * (t * t);
        const double fv = 0.0265;
        const double fa = 0.58 * fv;
        const double fmv = fa - fv, fpv = fa + fv, faM = fa * M;
        return prefactor * ((((2 * kl) * (((fmv * fmv) * kp) * pq - (M2 * (fa * fa + fv * fv)) * kq))
                             + m32 * ((faM * faM) * lq - (((2 * (fv * fv)) * lp) * pq)))
                            + (((2 * (fpv * fpv)) * kp) * kq) * lp);
    }
    case kM3VectorContact: {                                                          // :491-496
        const double c0 = K.c0;
        const double dd = ((c0 + 1) * m3) * (m32 - kp);
        const double denom = dd * dd;
        const double cMm = (c0 * M) * m3;
        const double numerator1 = (8 * m32) * ((((2 * c0) * (kq * lp)) * (kp - m32))
                                               - lq * (((((1 - c0 * c0) * (kp * kp)) - ((2 * m32) * kp)) + cMm * cMm) + m32 * m32));
        const double numerator2 = (16 * kl) * (((kq * ((c0 + 1) * kp + m3 * (c0 * M - m3))) * (m3 * (c0 * M + m3) - (c0 + 1) * kp))
                                               + ((c0 * m32) * pq) * (kp - m32));
        const double t = (((g * kGF) * K.fM) * K.ckm) / sqrt(2.0);
        return (-(t * t) * (numerator1 + numerator2)) / denom;
    }
    default:
        return 0.0;                                                                   // MatrixElementDecay3.__call__ :57-58
    }
}

// ---- QUADPACK dqk21 (21-point Gauss-Kronrod) -------------------------------------------------------------------------------
// xgk / wgk: Kronrod abscissae and weights (the even entries of xgk are the 10-point Gauss nodes); wg: Gauss weights.
#define ALP_GK21_XGK {0.995657163025808080735527280689003, 0.973906528517171720077964012084452, \
    0.930157491355708226001207180059508, 0.865063366688984510732096688423493, 0.780817726586416897063717578345042, \
    0.679409568299024406234327365114874, 0.562757134668604683339000099272694, 0.433395394129247190799265943165784, \
    0.294392862701460198131126603103866, 0.148874338981631210884826001129720, 0.0}
#define ALP_GK21_WGK {0.011694638867371874278064396062192, 0.032558162307964727478818972459390, \
    0.054755896574351996031381300244580, 0.075039674810919952767043140916190, 0.093125454583697605535065465083366, \
    0.109387158802297641899210590325805, 0.123491976262065851077958109831074, 0.134709217311473325928054001771707, \
    0.142775938577060080797094273138717, 0.147739104901338491374841515972068, 0.149445554002916905664936468389821}
#define ALP_GK21_WG {0.066671344308688137593568809893332, 0.149451349150580593145776339657697, \
    0.219086362515982043995534934228163, 0.269266719309996355091226921569469, 0.295524224714752870173892994651338}
#if defined(__CUDACC__)
static __constant__ double kGK21XgkDev[11] = ALP_GK21_XGK;
static __constant__ double kGK21WgkDev[11] = ALP_GK21_WGK;
static __constant__ double kGK21WgDev[5] = ALP_GK21_WG;
#endif
static const double kGK21XgkHost[11] = ALP_GK21_XGK;
static const double kGK21WgkHost[11] = ALP_GK21_WGK;
static const double kGK21WgHost[5] = ALP_GK21_WG;
#if defined(__CUDA_ARCH__)
#define ALP_XGK(i) kGK21XgkDev[i]
#define ALP_WGK(i) kGK21WgkDev[i]
#define ALP_WG(i) kGK21WgDev[i]
#else
#define ALP_XGK(i) kGK21XgkHost[i]
#define ALP_WGK(i) kGK21WgkHost[i]
#define ALP_WG(i) kGK21WgHost[i]
#endif

struct GK21 { double result, abserr, resabs, resasc; };

// dqk21 with the integrand f(x) = m2_meson3body(m122, x, m1, K); same evaluation and summation order as QUADPACK.
#if defined(__CUDACC__)
#define ALP_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ALP_HD_NOINLINE inline
#endif
// (not inlined on the device: the adaptive stage calls it from several places and each inlined copy carries the whole
// matrix-element switch)
ALP_HD_NOINLINE GK21 gk21_m2(double a, double b, double m122, double m1, const Meson3Const& K, double hv) {
    const double epmach = 2.220446049250313e-16, uflow = 2.2250738585072014e-308;
    const double centr = 0.5 * (a + b), hlgth = 0.5 * (b - a), dhlgth = fabs(hlgth);
    double fv1[10], fv2[10];
    double resg = 0.0;
    const double fc = m2_meson3body(m122, centr, m1, K, hv);
    double resk = ALP_WGK(10) * fc;
    double resabs = fabs(resk);
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {                         // Gauss nodes
        const int jtw = 2 * j + 1;
        const double absc = hlgth * ALP_XGK(jtw);
        const double f1 = m2_meson3body(m122, centr - absc, m1, K, hv), f2 = m2_meson3body(m122, centr + absc, m1, K, hv);
        fv1[jtw] = f1; fv2[jtw] = f2;
        const double fsum = f1 + f2;
        resg = resg + ALP_WG(j) * fsum;
        resk = resk + ALP_WGK(jtw) * fsum;
        resabs = resabs + ALP_WGK(jtw) * (fabs(f1) + fabs(f2));
    }
#pragma unroll 1
    for (int j = 0; j < 5; ++j) {                         // Kronrod nodes
        const int jtwm1 = 2 * j;
        const double absc = hlgth * ALP_XGK(jtwm1);
        const double f1 = m2_meson3body(m122, centr - absc, m1, K, hv), f2 = m2_meson3body(m122, centr + absc, m1, K, hv);
        fv1[jtwm1] = f1; fv2[jtwm1] = f2;
        const double fsum = f1 + f2;
        resk = resk + ALP_WGK(jtwm1) * fsum;
        resabs = resabs + ALP_WGK(jtwm1) * (fabs(f1) + fabs(f2));
    }
    const double reskh = resk * 0.5;
    double resasc = ALP_WGK(10) * fabs(fc - reskh);
#pragma unroll 1
    for (int j = 0; j < 10; ++j) resasc = resasc + ALP_WGK(j) * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    GK21 r;
    r.result = resk * hlgth;
    r.resabs = resabs * dhlgth;
    r.resasc = resasc * dhlgth;
    r.abserr = fabs((resk - resg) * hlgth);
    if (r.resasc != 0.0 && r.abserr != 0.0) {
        const double q = 200.0 * r.abserr / r.resasc;
        r.abserr = r.resasc * fmin(1.0, q * sqrt(q));     // (200 abserr/resasc)^1.5
    }
    if (r.resabs > uflow / (50.0 * epmach)) r.abserr = fmax((epmach * 50.0) * r.resabs, r.abserr);
    return r;
}

// quad(f, a, b)[0] for the matrix-element integrand: QAGS's first evaluation + acceptance test (dqagse), then QAGS's
// bisection of the interval with the largest error estimate until the summed error meets the bound (no extrapolation).
constexpr int kQagsIntervals = 8;

ALP_HD double qags21_m2(double a, double b, double m122, double m1, const Meson3Const& K) {
    const double epsabs = 1.49e-8, epsrel = 1.49e-8, epmach = 2.220446049250313e-16;
    const double hv = m2_threshold_factor(m122, K);
    GK21 r = gk21_m2(a, b, m122, m1, K, hv);
    double errbnd = fmax(epsabs, epsrel * fabs(r.result));
    const bool roundoff = r.abserr <= 100.0 * epmach * r.resabs && r.abserr > errbnd;        // ier = 2
    if (roundoff || (r.abserr <= errbnd && r.abserr != r.resasc) || r.abserr == 0.0) return r.result;
    // adaptive stage: the 21-point estimate is not trusted (integral above ~1e-8, or a peak the rule does not resolve --
    // the collinear peak of the *_ib1 models with an electron).  First bisection without bookkeeping arrays: QAGS's usual exit.
    const double mid = 0.5 * (a + b);
    const GK21 h1 = gk21_m2(a, mid, m122, m1, K, hv), h2 = gk21_m2(mid, b, m122, m1, K, hv);
    double area = (r.result + (h1.result + h2.result)) - r.result;
    double errsum = (r.abserr + (h1.abserr + h2.abserr)) - r.abserr;
    errbnd = fmax(epsabs, epsrel * fabs(area));
    if (errsum <= errbnd) return h2.abserr > h1.abserr ? (h2.result + h1.result) : (h1.result + h2.result);
    double al[kQagsIntervals], bl[kQagsIntervals], rl[kQagsIntervals], el[kQagsIntervals];
    {
        const bool sw = h2.abserr > h1.abserr;             // larger error estimate in slot 0 (dqagse labels 20/30)
        al[0] = sw ? mid : a;   bl[0] = sw ? b : mid;   rl[0] = sw ? h2.result : h1.result;   el[0] = sw ? h2.abserr : h1.abserr;
        al[1] = sw ? a : mid;   bl[1] = sw ? mid : b;   rl[1] = sw ? h1.result : h2.result;   el[1] = sw ? h1.abserr : h2.abserr;
    }
    int last = 2;
    while (last < kQagsIntervals) {
        int mx = 0;
        for (int k = 1; k < last; ++k) if (el[k] > el[mx]) mx = k;
        const double a1 = al[mx], b1 = 0.5 * (al[mx] + bl[mx]), a2 = b1, b2 = bl[mx];
        const GK21 r1 = gk21_m2(a1, b1, m122, m1, K, hv), r2 = gk21_m2(a2, b2, m122, m1, K, hv);
        const double area12 = r1.result + r2.result, erro12 = r1.abserr + r2.abserr;
        errsum = (errsum + erro12) - el[mx];
        area = (area + area12) - rl[mx];
        if (r2.abserr > r1.abserr) {
            al[mx] = a2; bl[mx] = b2; rl[mx] = r2.result; el[mx] = r2.abserr;
            al[last] = a1; bl[last] = b1; rl[last] = r1.result; el[last] = r1.abserr;
        } else {
            bl[mx] = b1; rl[mx] = r1.result; el[mx] = r1.abserr;
            al[last] = a2; bl[last] = b2; rl[last] = r2.result; el[last] = r2.abserr;
        }
        ++last;
        errbnd = fmax(epsabs, epsrel * fabs(area));
        if (errsum <= errbnd) break;
    }
    double result = 0.0;
    for (int k = 0; k < last; ++k) result = result + rl[k];
    return result;
}

// dGammadEa (fluxes.py:639-661 / 841-863).  Only M_E and M_MU have a matrix element attached (:649-661).
ALP_HD double meson3_dgamma_dea(double ea, double ml, const Meson3Const& K) {
    const double mm = K.mm, ma = K.ma;
    const double m212 = (mm * mm + ma * ma) - (2 * mm) * ea;
    const double rt = sqrt(m212);
    const double e2star = (m212 - ml * ml) / (2 * rt);
    const double e3star = ((mm * mm - m212) - ma * ma) / (2 * rt);
    if (ma > e3star) return 0.0;
    if (ml != kMMu && ml != kME) return 0.0;
    const double s = e2star + e3star;
    const double p2 = sqrt(e2star * e2star), p3 = sqrt(e3star * e3star - ma * ma);
    const double m223max = s * s - (p2 - p3) * (p2 - p3);
    const double m223min = s * s - (p2 + p3) * (p2 + p3);
    return K.pref * qags21_m2(m223min, m223max, m212, ml, K);
}

// ---- per-(meson, lepton) row records ----------------------------------------------------------------------------------------
enum Meson3Row { MR_ML = 0, MR_EAMIN, MR_SPAN, MR_BETA, MR_BOOST, MR_WGT, MR_COSLO, MR_COSSPAN, MR_SAC, MR_COSTH, MR_SINTH,
                 MR_POS, MR_LIVE, MR_NP };

// Isotropic class (fluxes.py:877-902): p = meson momentum, wgt = its weight
ALP_HD void meson3_row_isotropic(double p, double wgt, double ml, const Meson3Const& K, double* row) {
    const double ea_min = K.ma;
    const double ea_max = ((K.mm * K.mm + K.ma * K.ma) - ml * ml) / (2 * K.mm);
    const double beta = p / sqrt(p * p + K.mm * K.mm);                                // :888
    const double boost = pow(1 - beta * beta, -0.5);                                  // :889
    row[MR_ML] = ml; row[MR_EAMIN] = ea_min; row[MR_SPAN] = ea_max - ea_min; row[MR_BETA] = beta; row[MR_BOOST] = boost;
    row[MR_WGT] = wgt * (ea_max - ea_min);                                            // pion_wgt*mc_vol  (:898)
    row[MR_COSLO] = -1.0; row[MR_COSSPAN] = 1.0 - (-1.0);                             // uniform(-1, 1)
    row[MR_SAC] = 0.0; row[MR_COSTH] = 1.0; row[MR_SINTH] = 0.0; row[MR_POS] = 0.0; row[MR_LIVE] = 1.0;
}

// one Isotropic sample: returns e_lab, writes the weight
ALP_HD double meson3_sample_isotropic(double u_e, double u_c, const double* row, const Meson3Const& K, double& w) {
    const double ea = row[MR_EAMIN] + row[MR_SPAN] * u_e;                             // :882
    const double mom = sqrt(ea * ea - K.ma * K.ma);
    const double c = row[MR_COSLO] + row[MR_COSSPAN] * u_c;                           // :884
    const double pz = mom * c;
    const double e_lab = row[MR_BOOST] * (ea + row[MR_BETA] * pz);                    // :890
    w = ((row[MR_WGT] * meson3_dgamma_dea(ea, row[MR_ML], K)) / K.total_width) / K.n_samples;   // :898-899
    return e_lab;
}

// Decay-in-flight class (fluxes.py:669-725): meson = [p, theta, wgt], decay position x
ALP_HD void meson3_row_decay(double p, double theta, double wgt, double ml, double x, double det_dist, double det_length,
                             const Meson3Const& K, double* row) {
    const double ea_min = K.ma;
    const double ea_max = ((K.mm * K.mm + K.ma * K.ma) - ml * ml) / (2 * K.mm);
    const double beta = p / sqrt(p * p + K.mm * K.mm);                                // :682
    const double boost = pow(1 - beta * beta, -0.5);
    const double sac = cos(atan(det_length / (det_dist - x)));                        // :686
    const double ang = boost * acos(sac);
    const double min_cm_cos = cos(ang < kPi ? ang : kPi);                             // :687  cos(min(.., pi))
    row[MR_ML] = ml; row[MR_EAMIN] = ea_min; row[MR_SPAN] = ea_max - ea_min; row[MR_BETA] = beta; row[MR_BOOST] = boost;
    row[MR_WGT] = wgt;
    row[MR_COSLO] = min_cm_cos; row[MR_COSSPAN] = 1 - min_cm_cos;                     // uniform(min_cm_cos, 1)
    row[MR_SAC] = sac; row[MR_COSTH] = cos(theta); row[MR_SINTH] = sin(theta); row[MR_POS] = x;
    row[MR_LIVE] = (x > 52.0) ? 0.0 : 1.0;                                            // :671
}

struct Meson3DecayOut { double e_lab, cos_lab, theta_z, flux; int accept; };

ALP_HD Meson3DecayOut meson3_sample_decay(double u_e, double u_c, double u_phi, const double* row, const Meson3Const& K,
                                          double energy_cut) {
    Meson3DecayOut o;
    const double ea = row[MR_EAMIN] + row[MR_SPAN] * u_e;                             // :689
    const double mom = sqrt(ea * ea - K.ma * K.ma);
    const double c = row[MR_COSLO] + row[MR_COSSPAN] * u_c;                           // :691
    const double pz = mom * c;
    const double beta = row[MR_BETA], boost = row[MR_BOOST];
    const double e_lab = boost * (ea + beta * pz);                                    // :695
    const double pz_lab = boost * (pz + beta * ea);
    const double plab = sqrt(e_lab * e_lab - K.ma * K.ma);
    const double theta_lab = acos(pz_lab / plab);                                     // :697
    const double phi = 0.0 + (kTwoPi - 0.0) * u_phi;                                  // :698
    const double ct = cos(theta_lab);
    o.theta_z = acos(ct * row[MR_COSTH] + (cos(phi) * sin(theta_lab)) * row[MR_SINTH]);   // :701
    const double mc_vol_lab = (((1.0 - row[MR_COSLO]) * boost) * beta) * plab;        // :705
    const double w = (((row[MR_WGT] * mc_vol_lab) * meson3_dgamma_dea(ea, row[MR_ML], K)) / K.total_width) / K.n_samples;
    o.e_lab = e_lab;
    o.cos_lab = ct;
    o.flux = w * heaviside(e_lab - energy_cut, 1.0);                                  // :721
    o.accept = heaviside(ct - row[MR_SAC], 0.0) != 0.0 ? 1 : 0;                       // :716
    return o;
}

}  // namespace alp
