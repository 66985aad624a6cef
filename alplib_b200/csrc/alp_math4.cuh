// Generic 2 -> 2 scattering Monte Carlo: matrix elements (matrix_element.py:629-734), the atomic elastic form factor
// (form_factors.py:54-64) and one Scatter2to2MC sample (cross_section_mc.py:193-240).  Host + device like the rest of the
// alp_math headers (tests/hostcheck exercises the same expressions on the CPU); op order follows the reference expression
// by expression, compiled without FMA contraction.
#pragma once
#include "alp_math.cuh"

namespace alp {

enum M2Model : int {
    kM2InversePrimakoff = 0,      // a N -> gamma N      matrix_element.py:671-686
    kM2Primakoff = 1,             // gamma N -> a N      :690-705
    kM2InverseCompton = 2,        // a e- -> gamma e-    :650-666
    kM2Compton = 3,               // gamma e- -> a e-    :629-645
    kM2AssociatedProduction = 4   // e+ e- -> a gamma    :709-734
};

// Per-model constants evaluated by the host as Python-float arithmetic (like the reference's scalar sub-expressions).
struct M2Const {
    int model;
    double ma2, ma4;              // ma**2, ma**4
    double mN2, mN4;              // mN**2, mN**4                    (Primakoff-type)
    double z;                     // atomic number as a double
    double ff_a2;                 // a**2, a = 184.15*2.718**-0.5*z**(-1/3)/M_E   (form_factors.py:63)
    double pref;                  // model prefactor incl. coupling_product**2 (see m2_eval)
    double me2, me4;              // M_E**2, M_E**4
    double tmax;                  // associated production: soft-photon cut-off  (:729-731, depends on s only)
};

// form_factors.py:61-64  AtomicElasticFF.__call__(q): power(z*(t*a**2)/(1 + t*a**2), 2), t = q**2
ALP_HD double atomic_elastic_ff2(double q, double z, double a2) {
    const double t = q * q;
    const double x = (z * (t * a2)) / (1 + t * a2);
    return x * x;
}

// Squared matrix element M2(s, t) of `c.model`.  c.pref:
//   (inverse) Compton: z*4*pi*ALPHA*coupling_product**2;  inverse Primakoff: 2*coupling_product**2;  Primakoff:
//   coupling_product**2;  associated production: (pi*ALPHA)*coupling_product**2
ALP_HD double m2_eval(const M2Const& c, double s, double t) {
    if (c.model == kM2InverseCompton || c.model == kM2Compton) {
        const double me2 = c.me2, me4 = c.me4;
        const double u = (((2 * me2) + c.ma2) - s) - t;                                      // :638 / :659
        const double common = (me4 + me2 * (((3 * c.ma2) - (2 * s)) - t));
        const double num_s = common + s * ((s + t) - c.ma2);
        const double ds = me2 - s, du = me2 - u;
        const double Ms2 = (c.pref * num_s) / (ds * ds);                                    // power(M_E**2 - s, 2)
        const double Mu2 = (c.pref * num_s) / (du * du);
        const double MuMt = (c.pref * (common - (c.ma2 - s) * (s + t))) / (ds * du);
        const double m = (Ms2 + Mu2) + 2 * MuMt;
        return c.model == kM2InverseCompton ? 2 * m : m;
    }
    if (c.model == kM2InversePrimakoff || c.model == kM2Primakoff) {
        const double prop = t * t;                                                          // power(t, 2)
        // -t*(2*mN**2*(ma**2 - 2*s - t) + 2*mN**4 - 2*ma**2*(s + t) + ma**4 + 2*s**2 + 2*s*t + t**2)
        const double poly = ((((((2 * c.mN2) * ((c.ma2 - 2 * s) - t) + 2 * c.mN4) - (2 * c.ma2) * (s + t)) + c.ma4)
                              + 2 * (s * s)) + (2 * s) * t) + t * t;
        const double num = (-t) * poly;
        return ((atomic_elastic_ff2(sqrt(fabs(t)), c.z, c.ff_a2) * c.pref) * num) / prop;
    }
    // associated production e+ e- -> a gamma (same expressions as pairann_sample / M2AssociatedProduction :717-734)
    const double me2 = c.me2, me4 = c.me4;
    const double u_prop = ((me2 + c.ma2) - s) - t;
    const double t_prop = me2 - t;
    const double tt = t * (((-c.ma2) + s) + t);
    const double Mt2 = (-4 * (((3 * me4) - me2 * (c.ma2 + s)) + tt)) / (t_prop * t_prop);
    const double Mu2 = (-4 * (((7 * me4) + me2 * ((c.ma2 - 3 * s) - 4 * t)) + tt)) / (u_prop * u_prop);
    const double MtMu = (8 * ((((-3) * me4) + me2 * (s - 2 * t)) + tt)) / (u_prop * t_prop);
    const double M2 = (Mt2 + Mu2) + 2 * MtMu;
    return (c.pref * M2) * heaviside(c.tmax - t, 0.0);
}

// Per-call constants of Scatter2to2MC.scatter_sim, evaluated by the host exactly as the reference evaluates them
// (cross_section_mc.py:207-226): everything that does not depend on the variates.
struct Scatter2Const {
    M2Const m2;
    double s;                 // cm_p4.mass2()
    double msum;              // m1**2 + m3**2
    double pp;                // p1_cm*p3_cm
    double ee;                // e1_cm*e3_cm
    double p3_cm, e3_cm;
    double vol;               // linear: 4*p1_cm*p3_cm/n_samples;  log sampling: unused
    double p1_cm, n_samples;  // log sampling: exp(logu)*2*p1_cm*p3_cm/n_samples, left to right
    double den;               // 16*pi*(s - (m1+m2)**2)*(s - (m1-m2)**2)
    double mat[16];           // lorentz_boost matrix for frame velocity -v_in (identity when beta == 0)
    double p12[4];            // lv_p1 + lv_p2
    int log_sampling;
};

struct Scatter2Out { double t, w, p3cm[4], p3lab[4], p4lab[4], cosine; };

// One sample: u1 -> phi, u2 -> theta (cross_section_mc.py:199-204, 220-238)
ALP_HD void scatter2_sample(double u1, double u2, const Scatter2Const& c, Scatter2Out& o) {
    const double phi = kTwoPi * u1;
    double th, vol;
    if (c.log_sampling) {
        const double logu = -12 + 12 * u2;                                         // np.random.uniform(-12, 0): lo + (hi-lo)*u
        const double eu = exp(logu);
        th = acos(1 - 2 * eu);                                                     // :201
        vol = (((eu * 2) * c.p1_cm) * c.p3_cm) / c.n_samples;                      // :224
    } else {
        th = acos(1 - 2 * u2);                                                     // :204
        vol = c.vol;                                                               // :226
    }
    const double cth = cos(th), sth = sin(th);
    o.t = c.msum + 2 * (c.pp * cth - c.ee);                                        // :220
    o.w = vol * (m2_eval(c.m2, c.s, o.t) / c.den);                                 // :227, :184-185
    o.p3cm[0] = c.e3_cm;
    o.p3cm[1] = (c.p3_cm * cos(phi)) * sth;                                        // :230-233
    o.p3cm[2] = (c.p3_cm * sin(phi)) * sth;
    o.p3cm[3] = c.p3_cm * cth;
    boost_apply(c.mat, o.p3cm[0], o.p3cm[1], o.p3cm[2], o.p3cm[3], o.p3lab);       // :234
#pragma unroll
    for (int k = 0; k < 4; ++k) o.p4lab[k] = c.p12[k] - o.p3lab[k];                // :238
    const double pm = sqrt((o.p3lab[1] * o.p3lab[1] + o.p3lab[2] * o.p3lab[2]) + o.p3lab[3] * o.p3lab[3]);
    o.cosine = o.p3lab[3] / pm;                                                    // LorentzVector.cosine :119-120
}

}  // namespace alp
