// Per-row / per-sample arithmetic of the e+- channels: bremsstrahlung with track-length attenuation
// (fluxes.py:257-356, prod_xs.py:209-221, 313-323), resonant production (fluxes.py:421-445) and associated production
// e+ e- -> a gamma (fluxes.py:503-518, cross_section_mc.py:192-240, matrix_element.py:717-734).
// Same conventions as alp_math.cuh: host+device, reference op order, no FMA contraction.
#pragma once
#include "alp_math.cuh"

namespace alp {

constexpr int kGL = 32;     // Gauss-Legendre points per panel of the track-length integral

// ---- fmath.py:33-36 nu_integral(x, alpha=-1, T_max) = int_0^T x^(t-1)/Gamma(t) dt ---------------------------------------
// The reference calls QUADPACK (scipy.integrate.quad); its result is exact to ~2e-16 on the range reached here.  A
// 32-point Gauss-Legendre rule on the panels [0,1] and [1,T] agrees with it to <= 2e-15 for x in [1e-8, 13]
// (measured against mpmath).  gl_x/gl_w: nodes/weights on [-1,1] (numpy.polynomial.legendre.leggauss(32)).
ALP_HD double nu_integral_m1(double x, double T, const double* gl_x, const double* gl_w) {
    const double lx = log(x);
    double total = 0.0;
    const int panels = T > 1.0 ? 2 : 1;
    for (int p = 0; p < panels; ++p) {
        const double a = p == 0 ? 0.0 : 1.0;
        const double b = (panels == 1 || p == 1) ? T : 1.0;
        const double h = 0.5 * (b - a), c = 0.5 * (b + a);
        double acc = 0.0;
        for (int i = 0; i < kGL; ++i) {
            const double t = h * gl_x[i] + c;
            acc += gl_w[i] * exp((t - 1.0) * lx - lgamma(t));
        }
        total += h * acc;
    }
    return total;
}

// fluxes.py:264-266
ALP_HD double track_length_integrated_prob(double ei, double ef, double t_arg, const double* gl_x, const double* gl_w) {
    const double h = heaviside(ei - ef, 1.0);
    if (h == 0.0) return 0.0;       // (3/4)*0*nu/Ei; avoids log of a ratio < 1 raised to a real power
    return ((0.75 * h) * nu_integral_m1(log(ei / ef), t_arg, gl_x, gl_w)) / ei;
}

// fluxes.py:257-259
ALP_HD double track_length_prob(double ei, double ef, double t) {
    const double b = 4.0 / 3.0;
    return heaviside(ei - ef, 1.0) * fabs(pow(log(ei / ef), b * t - 1) / (ei * tgamma(b * t)));
}

// ---- bremsstrahlung ------------------------------------------------------------------------------------------------------
enum BremRow { BR_E1 = 0, BR_W, BR_DL, BR_PREF, BR_EMIT, BR_DLX, BR_NP };

struct BremGlobals {
    double ma, ma2;
    double log10_ma;        // np.log10(ma)
    double ep_min;          // max(ma, M_E)
    double nt;              // ntarget_area_density * HBARC**2
    double pref_num;        // 2*r0**2*g**2/4/pi          (pseudoscalar, divided by E1 per row)
    double zl;              // z**2*ln_el + z*ln_inel
    double zz;              // z**2 + z
    double pref_vec;        // chi*(4*ALPHA**2*g**2)/(4*pi) (vector)
    double ln10;            // np.log(10)
    double n_samples;
    double t_arg;           // 4*max_t/3
    int vector;             // boson_type == "vector"
    int use_track_length;
};

// Row prelude.  e0/w0: the electron row; u: its scalar uniform (track-length mode).  fl = Phi_e(E0)+Phi_p(E0).
ALP_HD void brem_row_params(double e0, double w0, double u, double fl, const BremGlobals& g, const double* gl_x,
                            const double* gl_w, double* out) {
    double e1, w;
    if (g.use_track_length) {
        e1 = g.ep_min + (e0 - g.ep_min) * u;                                     // fluxes.py:350
        w = (e0 - g.ep_min) * (fl * track_length_integrated_prob(e0, e1, g.t_arg, gl_x, gl_w));   // :351, :314-316
    } else {
        e1 = e0; w = w0;                                                         // :354-356
    }
    const double r = g.ma / e1;
    const double ea_max = e1 * (1 - r * r);                                      // :322
    const bool emit = ea_max > g.ma;                                             // :323
    out[BR_E1] = e1;
    out[BR_W] = w;
    out[BR_DL] = emit ? (log10(ea_max) - g.log10_ma) : 0.0;                      // :326, :329
    out[BR_PREF] = g.pref_num / e1;                                              // prod_xs.py:217
    out[BR_EMIT] = emit ? 1.0 : 0.0;
    out[BR_DLX] = emit ? (log10(ea_max / e1) - log10(g.ma / e1)) : 0.0;          // fluxes.py:333
}

// prod_xs.py:209-221 with the z-only logs hoisted
ALP_HD double brem_dsigma_dea(double ea, double e1, double pref_row, const BremGlobals& g) {
    const double x = ea / e1;
    const double q = g.ma / (x * kME);
    const double f = (q * q) * (1 - x);
    const double opf = 1 + f;
    const double opf2 = opf * opf;
    const double t1 = ((x * (1 + f / 1.5)) / opf2) * g.zl;
    const double a = (opf * log(opf)) / (3 * (f * f));
    const double b = ((1 + 4 * f) + 2 * (f * f)) / ((3 * f) * opf2);
    const double ps = t1 + (x * g.zz) * (a - b);
    return (pref_row * ps) * heaviside(ps, 0.0);
}

// prod_xs.py:313-323
ALP_HD double brem_dsigma_dx_vector(double x, const BremGlobals& g) {
    const double me2 = kME * kME;
    return (g.pref_vec * ((1 - x) + (x * x) / 3)) / (((g.ma2 * (1 - x)) / x) + x * me2);
}

// fluxes.py:326-337: one sample of an emitting row.  Returns E_a, writes the production weight.
ALP_HD double brem_sample(double u, const double* b, const BremGlobals& g, double& flux) {
    const double lg = g.log10_ma + b[BR_DL] * u;                                 // uniform(log10 ma, log10 ea_max)
    const double ea = pow(10.0, lg);
    double diff_br;
    if (!g.vector) {
        const double mc_vol = ((g.ln10 * ea) * b[BR_DL]) / g.n_samples;          // :329
        diff_br = (g.nt * mc_vol) * brem_dsigma_dea(ea, b[BR_E1], b[BR_PREF], g);   // :330
    } else {
        const double x = ea / b[BR_E1];                                          // :332
        const double mc_vol = ((g.ln10 * x) * b[BR_DLX]) / g.n_samples;          // :333
        diff_br = (g.nt * mc_vol) * brem_dsigma_dx_vector(x, g);                 // :334
    }
    flux = b[BR_W] * diff_br;                                                    // :337
    return ea;
}

// Fast mode (<= 1e-9 relative): 10^x from the exp table instead of pow (E_a to ~1 ulp), the eight divisions of the spectrum
// folded into two reciprocals (1/(x m_e) and 1/(3 f^2 (1+f)^2), from which 1/(1+f)^2, 1/(3f^2) and 1/(3f(1+f)^2) follow).
// Near the kinematic end point (f < 1e-3) the bracket a - b cancels like 1/f and the reference expression is used.
ALP_HD double brem_sample_fast(double u, const double* b, const BremGlobals& g, double& flux, const double* exp_tab = nullptr) {
    const double lg = g.log10_ma + b[BR_DL] * u;
    if (!(fabs(lg) <= 300.0)) return brem_sample(u, b, g, flux);
    const double ea = exp10_fast_core(lg, exp_tab);
    const double x = ea * rcp_fast(b[BR_E1]);
    double diff_br;
    if (!g.vector) {
        const double q = g.ma * rcp_fast(x * kME);
        const double omx = 1 - x;
        const double f = (q * q) * omx;
        if (!(f > 1e-3) || !(f < 1e100)) return brem_sample(u, b, g, flux);
        const double opf = 1 + f, opf2 = opf * opf, f2 = f * f;
        const double R = rcp_fast((3 * f2) * opf2);
        const double t1 = ((x * (1 + f * (1.0 / 1.5))) * ((3 * f2) * R)) * g.zl;
        const double aa = (opf * log(opf)) * (opf2 * R);
        const double bb = ((1 + 4 * f) + 2 * f2) * (f * R);
        const double ps = t1 + (x * g.zz) * (aa - bb);
        const double mc_vol = ((g.ln10 * ea) * b[BR_DL]) / g.n_samples;
        diff_br = (g.nt * mc_vol) * (ps > 0.0 ? b[BR_PREF] * ps : (ps == ps ? 0.0 * (b[BR_PREF] * ps) : ps));
    } else {
        const double omx = 1 - x;
        const double den = g.ma2 * omx + (x * x) * (kME * kME);     // x * ( ma2 (1-x)/x + x me2 )
        const double mc_vol = ((g.ln10 * x) * b[BR_DLX]) / g.n_samples;
        diff_br = (g.nt * mc_vol) * (((g.pref_vec * (omx + (x * x) / 3)) * x) * rcp_fast(den));
    }
    flux = b[BR_W] * diff_br;
    return ea;
}

// ---- resonance: one term of the (E+, t) Monte-Carlo integral (fluxes.py:437-441, 410-411) -------------------------------------
ALP_HD double resonance_term(double u1, double u2, double eres, double emax, double max_t, const double* pos_e,
                             const double* pos_f, int n_pos) {
    const double e = eres + (emax - eres) * u1;                                  // :437
    const double t = 0.0 + (max_t - 0.0) * u2;                                   // :438
    return interp(e, pos_e, pos_f, n_pos, 0.0, 0.0) * track_length_prob(e, eres, t);
}

// ---- fast mode of the resonance term (<= 1e-9 relative) ----------------------------------------------------------------------
// 1/Gamma(z) for 0 < z < 170: recurrence to w = z - k in [1, 2), degree-14 polynomial in (w - 1.5) (Chebyshev-node
// interpolant, 1.7e-16 relative), one reciprocal for the product (z-1)...(z-k).
ALP_HD double rgamma_fast(double z) {
    const double c[15] = {1.1283791670955126, -0.04117452644528216, -0.5266544355255445, 0.17510202604379396,
                          0.05096686024771327, -0.04215516936245979, 0.006612897826512947, 0.002120731326787741,
                          -0.00111073024865264, 0.00015235875380774502, 2.5355146572834772e-05, -1.3902747279703676e-05,
                          2.1565111610410453e-06, 7.383394654158106e-08, -8.996935696824659e-08};
    double w = z, scale = 1.0;
    bool inv = false;
    if (z < 1.0) { scale = z; w = z + 1.0; }               // 1/Gamma(z) = z / Gamma(z + 1)
    else {
        double prod = 1.0;
        while (w >= 2.0) { w -= 1.0; prod *= w; inv = true; }      // Gamma(z) = (z-1)(z-2)...(w) Gamma(w)
        scale = prod;
    }
    const double x = w - 1.5;
    double p = c[14];
#pragma unroll
    for (int k = 13; k >= 0; --k) p = fma(p, x, c[k]);
    return inv ? p * rcp_fast(scale) : p * scale;
}

// np.interp on a uniformly spaced abscissa: index from the spacing, verified against the table (falls back to the
// binary search of interp() when the guess is off by more than one or the table is not uniform: inv_dx = 0)
ALP_HD double interp_uniform(double x, const double* xp, const double* fp, int n, double left, double right, double inv_dx) {
    if (!(inv_dx > 0.0) || n < 3 || !(x > xp[0]) || !(x < xp[n - 1])) return interp(x, xp, fp, n, left, right);
    int i = (int)((x - xp[0]) * inv_dx);
    i = i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
    if (x < xp[i]) { if (i > 0 && x >= xp[i - 1]) i = i - 1; else return interp(x, xp, fp, n, left, right); }
    else if (!(x < xp[i + 1])) { if (i + 2 <= n - 1 && x < xp[i + 2]) i = i + 1; else return interp(x, xp, fp, n, left, right); }
    if (xp[i] == x) return fp[i];
    const double slope = (fp[i + 1] - fp[i]) / (xp[i + 1] - xp[i]);
    double r = slope * (x - xp[i]) + fp[i];
    if (r != r) return interp(x, xp, fp, n, left, right);
    return r;
}

// track_length_prob with pow(L, y) = exp(y log L) from the exp table and 1/Gamma from rgamma_fast (fluxes.py:257-259)
ALP_HD double resonance_term_fast(double u1, double u2, double eres, double emax, double max_t, const double* pos_e,
                                  const double* pos_f, int n_pos, double inv_dx, const double* exp_tab = nullptr) {
    const double e = eres + (emax - eres) * u1;
    const double t = 0.0 + (max_t - 0.0) * u2;
    const double z = (4.0 / 3.0) * t;
    const double L = log(e / eres);
    const double y = (z - 1) * log(L);
    if (!(L > 0.0) || !(fabs(y) <= 700.0) || !(z > 1e-300) || !(z < 170.0))
        return resonance_term(u1, u2, eres, emax, max_t, pos_e, pos_f, n_pos);
    const double tl = fabs((exp_fast_core(y, exp_tab) * rgamma_fast(z)) * rcp_fast(e));
    return interp_uniform(e, pos_e, pos_f, n_pos, 0.0, 0.0, inv_dx) * tl;
}

// ---- associated production e+ e- -> a gamma --------------------------------------------------------------------------------------
enum PairRow { PA_S = 0, PA_P1P3, PA_E1E3, PA_E3, PA_P3, PA_GAM, PA_M03, PA_VOL, PA_DEN, PA_TMAX, PA_C, PA_NP };

struct PairGlobals {
    double ma, ma2;
    double me2, me4;        // M_E**2, M_E**4 (Python float pow)
    double pref;            // (pi*ALPHA) * 1.0**2
    double wconst;          // z * ge**2 * ntarget_area_density * HBARC**2 applied left to right after pos_wgt
    double z, ge2, ntad, hbarc2;
    double n_samples;
};

// Per positron row (already above threshold, fluxes.py:507): CM kinematics and the boost back to the lab.
ALP_HD void pairann_row_params(double ep, double wgt, const PairGlobals& g, double* out) {
    const double m1 = kME, m2 = kME, m4 = 0.0;
    const double p1z = sqrt(ep * ep - g.me2);                                    // fluxes.py:513
    const double e_in = ep + kME;                                                // cross_section_mc.py:208-209
    const double pz_in = p1z + 0.0;
    const double vz = pz_in / e_in;                                              // :211
    const double s = ((e_in * e_in - 0.0) - 0.0) - pz_in * pz_in;                // mass2 :110-111
    const double a1 = (s - m1 * m1) - m2 * m2;
    const double b1 = (2 * m1) * m2;
    const double p1 = sqrt((a1 * a1 - b1 * b1) / (4 * s));                       // :186
    const double a3 = (s - g.ma2) - m4 * m4;
    const double b3 = (2 * g.ma) * m4;
    const double p3 = sqrt((a3 * a3 - b3 * b3) / (4 * s));                       // :189
    const double e1 = sqrt(p1 * p1 + m1 * m1);                                   // :218
    const double e3 = sqrt(p3 * p3 + g.ma2);                                   // :219
    // lorentz_boost(p3, -v_in): v = (0,0,-vz)  (:235, :676-685)
    const double nv = -vz;
    const double beta = sqrt(nv * nv);
    double gam = 1.0, m03 = 0.0;
    if (beta != 0.0) {
        const double nz = nv / beta;
        gam = 1 / sqrt(1 - beta * beta);
        m03 = ((-gam) * beta) * nz;
    }
    // soft-photon cut (matrix_element.py:730-732)
    const double sa = s - 2 * g.me2;
    const double ppos = sqrt((sa * sa - 4 * g.me4) / (4 * s));
    const double epos = sqrt(ppos * ppos + g.me2);
    out[PA_S] = s;
    out[PA_P1P3] = p1 * p3;
    out[PA_E1E3] = e1 * e3;
    out[PA_E3] = e3;
    out[PA_P3] = p3;
    out[PA_GAM] = gam;
    out[PA_M03] = m03;
    out[PA_VOL] = ((4 * p1) * p3) / g.n_samples;                                 // :227
    const double sp = m1 + m2, sm = m1 - m2;
    out[PA_DEN] = ((16 * kPi) * (s - sp * sp)) * (s - sm * sm);                  // :181-182
    out[PA_TMAX] = g.me2 - ((2 * 0.01) * epos) * epos;                           // matrix_element.py:732
    out[PA_C] = (((wgt * g.z) * g.ge2) * g.ntad) * g.hbarc2;                     // fluxes.py:518
}

// One sample: u -> theta (the phi draw does not enter the lab energy for a beam along z).
ALP_HD double pairann_sample(double u, const double* r, const PairGlobals& g, double& flux) {
    const double theta = acos(1 - 2 * u);                                        // cross_section_mc.py:205
    const double cth = cos(theta);
    const double s = r[PA_S];
    const double m1sq = g.me2, m3sq = g.ma2;
    const double t = (m1sq + m3sq) + 2 * (r[PA_P1P3] * cth - r[PA_E1E3]);        // :221
    // matrix_element.py:717-734
    const double u_prop = ((g.me2 + g.ma2) - s) - t;
    const double t_prop = g.me2 - t;
    const double tail = t * (((-g.ma2) + s) + t);
    const double mt2 = (-4 * ((3 * g.me4 - g.me2 * (g.ma2 + s)) + tail)) / (t_prop * t_prop);
    const double mu2 = (-4 * ((7 * g.me4 + g.me2 * ((g.ma2 - 3 * s) - 4 * t)) + tail)) / (u_prop * u_prop);
    const double mtmu = (8 * (((-3) * g.me4 + g.me2 * (s - 2 * t)) + tail)) / (u_prop * t_prop);
    const double m2 = (mt2 + mu2) + 2 * mtmu;
    const double val = (g.pref * m2) * heaviside(r[PA_TMAX] - t, 0.0);
    const double wts = r[PA_VOL] * (val / r[PA_DEN]);                            // :228
    flux = r[PA_C] * wts;
    const double pz = r[PA_P3] * cth;                                            // :234
    return r[PA_GAM] * r[PA_E3] + r[PA_M03] * pz;                                // boost row 0
}

// Fast mode (<= 1e-9 relative): cos(acos(1 - 2u)) = 1 - 2u, and the three propagator divisions share one reciprocal
// 1/(t_prop u_prop).
ALP_HD double pairann_sample_fast(double u, const double* r, const PairGlobals& g, double& flux) {
    const double cth = 1 - 2 * u;
    const double s = r[PA_S];
    const double t = (g.me2 + g.ma2) + 2 * (r[PA_P1P3] * cth - r[PA_E1E3]);
    const double u_prop = ((g.me2 + g.ma2) - s) - t;
    const double t_prop = g.me2 - t;
    const double den = t_prop * u_prop;
    if (!(fabs(den) > 1e-290) || !(fabs(den) < 1e290)) return pairann_sample(u, r, g, flux);
    const double R = rcp_fast(den);
    const double it = u_prop * R, iu = t_prop * R;                 // 1/t_prop, 1/u_prop
    const double tail = t * (((-g.ma2) + s) + t);
    const double mt2 = (-4 * ((3 * g.me4 - g.me2 * (g.ma2 + s)) + tail)) * (it * it);
    const double mu2 = (-4 * ((7 * g.me4 + g.me2 * ((g.ma2 - 3 * s) - 4 * t)) + tail)) * (iu * iu);
    const double mtmu = (8 * (((-3) * g.me4 + g.me2 * (s - 2 * t)) + tail)) * R;
    const double m2 = (mt2 + mu2) + 2 * mtmu;
    const double val = (g.pref * m2) * heaviside(r[PA_TMAX] - t, 0.0);
    const double wts = r[PA_VOL] * (val * rcp_fast(r[PA_DEN]));
    flux = r[PA_C] * wts;
    return r[PA_GAM] * r[PA_E3] + r[PA_M03] * (r[PA_P3] * cth);
}

// ---- e+ e- -> gamma a through a virtual photon (fluxes.py:1058-1131, prod_xs.py:109-129) ----------------------------------
struct PairGammaGlobals {
    double ma, ma2, ma4;
    double me2, me4;          // M_E**2, M_E**4
    double ep_min;            // max((ma**2 - M_E**2)/(2*M_E), M_E)
    double nt;                // ntarget_area_density * HBARC**2
    double zc;                // z*4*pi*ALPHA*g**2  (left to right)
    double n_samples;
    double max_t;             // 5.0 radiation lengths (fluxes.py:1099,1103)
    int fast;                 // fast arithmetic mode of the track-length factor (exp-table pow, polynomial 1/Gamma)
};

// fluxes.py:1078-1080 (note heaviside(., 0.0) here, 1.0 in the module-level track_length_prob)
ALP_HD double track_length_prob0(double ei, double ef, double t) {
    const double b = 4.0 / 3.0;
    return heaviside(ei - ef, 0.0) * fabs(pow(log(ei / ef), b * t - 1) / (ei * tgamma(b * t)));
}

// Fast mode of the same factor (<= 1e-9 relative): pow(L, y) = exp(y log L) from the exp table, 1/Gamma from rgamma_fast.
ALP_HD double track_length_prob0_fast(double ei, double ef, double t) {
    const double z = (4.0 / 3.0) * t;
    const double L = log(ei / ef);
    const double y = (z - 1) * log(L);
    if (!(ei > ef) || !(L > 0.0) || !(fabs(y) <= 700.0) || !(z > 1e-300) || !(z < 170.0)) return track_length_prob0(ei, ef, t);
    return fabs((exp_fast_core(y) * rgamma_fast(z)) * rcp_fast(ei));
}

// prod_xs.py:109-129
ALP_HD double epem_to_alp_photon_dsigma_de(double ea, double ep, const PairGammaGlobals& g) {
    const double s = (2 * kME) * (kME + ep);
    const double d = s - g.ma2;
    const double ps = sqrt((d * d) / (4 * s));
    const double a = s - 2 * g.me2;
    const double pp = sqrt((a * a - 4 * g.me4) / (4 * s));
    const double es = sqrt(ps * ps + g.ma2);
    const double epc = sqrt(pp * pp + g.me2);
    const double beta = sqrt(ep * ep - g.me2) / (kME + ep);
    const double gam = pow(1 - beta * beta, -0.5);
    const double ct = (ea / gam - es) / (beta * ps);
    const double t = (g.ma2 + g.me2) - 2 * (epc * es - (pp * ps) * ct);
    const double inner = ((2 * s) * g.me4 + (2 * g.me2) * ((g.ma4 - s * g.ma2) - (2 * s) * t))
                         + s * (((((g.ma4 - (2 * g.ma2) * (s + t)) + s * s) + (2 * s) * t)) + 2 * (t * t));
    const double m_st = (g.zc * inner) / (s * s);
    const double jac = ((2 * pp) / gam) / beta;
    return ((heaviside(ep - g.ep_min, 1.0) * jac) * m_st) / (((16 * kPi) * (s - 4 * g.me2)) * s);
}

// One sample of bin (center e_lab, width, dN/dE at the centre): u1 -> depth t, u2 -> smeared E+, u3 -> E_a.
ALP_HD double pairgamma_sample(double u1, double u2, double u3, double e_lab, double width, double flux_c,
                               const PairGammaGlobals& g, double& flux) {
    const double t = 0.0 + (g.max_t - 0.0) * u1;                                 // fluxes.py:1099
    const double ep = g.ep_min + (e_lab - g.ep_min) * u2;                        // :1100
    const double fw = flux_c * (g.fast ? track_length_prob0_fast(e_lab, ep, t) : track_length_prob0(e_lab, ep, t));   // :1102, :1085-1086
    const double pos_flux = (((width * 5.0) * (e_lab - g.ep_min)) * fw) / g.n_samples;   // :1103
    const double beta = sqrt(ep * ep - g.me2) / (kME + ep);                      // :1118
    const double gam = pow(1 - beta * beta, -0.5);
    const double s = (2 * kME) * (kME + ep);
    const double d = s - g.ma2;
    const double ps = sqrt((d * d) / (4 * s));
    const double es = sqrt(ps * ps + g.ma2);
    const double es_min = gam * (es - ps * beta);                                // :1124
    const double es_max = gam * (es + ps * beta);
    const double ea = es_min + (es_max - es_min) * u3;                           // :1126
    const double mc_vol = ((2 * gam) * beta) * ps;                               // :1127
    const double cm = g.nt * epem_to_alp_photon_dsigma_de(ea, ep, g);            // :1129-1130
    flux = (cm * pos_flux) * mc_vol;                                             // :1134
    return ea;
}

}  // namespace alp
