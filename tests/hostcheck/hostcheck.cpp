// TEST INFRASTRUCTURE ONLY -- host build of the kernels' per-sample arithmetic (alplib_b200/csrc/alp_math.cuh).
//
// Compiled by tests/conftest.py with `g++ -O2 -ffp-contract=off` so the CPU-only test tier can check the exact
// expressions the CUDA kernels evaluate (same header, glibc libm instead of libdevice) against the golden fixtures
// without a GPU.  Nothing under alplib_b200/ loads this library; the product path is the CUDA kernels only.
#include <stdint.h>

#include "../../alplib_b200/csrc/alp_math.cuh"
#include "../../alplib_b200/csrc/alp_math2.cuh"
#include "../../alplib_b200/csrc/alp_math3.cuh"
#include "../../alplib_b200/csrc/alp_math4.cuh"
#include "../../include/alpk.h"

using namespace alp;

static PropConst to_prop(const alpk_prop_t* p) {
    PropConst c;
    c.ma = p->ma; c.ma2 = p->ma2; c.width = p->width; c.rescale = p->rescale;
    c.neg_dist_mbm = p->neg_dist_mbm; c.neg_len_mbm = p->neg_len_mbm; c.geom = p->geom; c.geom32 = p->geom32;
    c.apply_geom = p->apply_geom; c.geom_f64 = p->geom_f64; c.use_tof = p->use_tof; c.tof_light = p->tof_light;
    c.det_dist = p->det_dist; c.timing_window = p->timing_window;
    c.fast = p->fast; c.k1 = p->k1; c.k2 = p->k2;
    prop_derive(c);
    return c;
}

static EventConst to_event(const alpk_event_t* e) {
    EventConst c;
    c.days_sec = e->days_sec; c.days_sec32 = e->days_sec32; c.days_f64 = e->days_f64; c.threshold = e->threshold;
    return c;
}

extern "C" {

void hc_philox(uint64_t seed, uint32_t tag, uint32_t row, uint64_t n, int two, double* out) {
    for (uint64_t s = 0; s < n; ++s) {
        if (two) uniform_2(seed, tag, row, s, out[2 * s], out[2 * s + 1]);
        else out[s] = uniform_1(seed, tag, row, s);
    }
}

void hc_philox_rounds(uint64_t seed, uint32_t tag, uint32_t row, uint64_t n, int rounds, double* out) {
    for (uint64_t s = 0; s < n; ++s) out[s] = uniform_1(seed, tag, row, s, rounds);
}

void hc_exp_fast(const double* x, uint64_t n, double* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = exp_fast(x[i]);
}

void hc_abs_sigma(const double* le, const double* lx, int nt, const double* e, uint64_t n, double* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = abs_sigma_mev(e[i], le, lx, nt);
}

void hc_primakoff(const double* eg, const double* phi, uint64_t n, const double* le, const double* lx, int nt,
                  double ma, double pref, double r02, double* oe, double* of) {
    for (uint64_t i = 0; i < n; ++i) {
        oe[i] = eg[i];
        of[i] = phi[i] * (primakoff_sigma(eg[i], ma, pref, r02) / abs_sigma_mev(eg[i], le, lx, nt));
    }
}

void hc_compton(const double* eg, const double* phi, uint64_t n_rows, uint32_t n_samples, const double* le,
                const double* lx, int nt, const alpk_compton_t* par, const double* u, double* oe, double* of) {
    ComptonGlobals g{par->ma, par->ma2, par->aa, par->z, par->n_samples};
    for (uint64_t r = 0; r < n_rows; ++r) {
        double row[CB_NP];
        compton_bin_params(eg[r], phi[r], abs_sigma_mev(eg[r], le, lx, nt), g, row);
        for (uint32_t s = 0; s < n_samples; ++s) {
            const uint64_t i = r * n_samples + s;
            oe[i] = par->fast ? compton_sample_fast(u[i], row, g, of[i]) : compton_sample(u[i], row, g, of[i]);
        }
    }
}

void hc_propagate(const double* e, const double* w, uint64_t n, const alpk_prop_t* p, float* d32, float* s32,
                  double* d64, double* s64) {
    const PropConst c = to_prop(p);
    for (uint64_t i = 0; i < n; ++i) {
        if (c.fast) propagate_sample_fast(e[i], w[i], c, d64[i], s64[i]);
        else propagate_sample(e[i], w[i], c, d64[i], s64[i]);
        d32[i] = prop_cast(d64[i], c);
        s32[i] = prop_cast(s64[i], c);
    }
}

void hc_decays(const double* e, const float* w32, uint64_t n, const alpk_event_t* ev, double* out) {
    const EventConst c = to_event(ev);
    for (uint64_t i = 0; i < n; ++i) out[i] = decay_event_weight(e[i], w32[i], c);
}

void hc_scatter(int kind, const double* e, const float* w32, uint64_t n, const double* xs, double prefactor,
                double mbm2, const alpk_event_t* ev, double* out) {
    const EventConst c = to_event(ev);
    IComptonConst ic{xs[0], xs[1], xs[2], xs[3], xs[4], xs[5]};
    const double lead = c.days_sec * prefactor;
    for (uint64_t i = 0; i < n; ++i) {
        double s;
        if (kind == 0) s = icompton_sigma(e[i], ic);
        else if (kind == 1) s = iprimakoff_sigma(e[i], xs[0], xs[1], xs[2], xs[3]);
        else if (kind == 2) s = icompton_sigma_fast(e[i], ic);
        else s = iprimakoff_sigma_fast(e[i], xs[0], xs[1], xs[2], xs[3]);
        out[i] = (((lead * s) * mbm2) * (double)w32[i]) * heaviside(e[i] - c.threshold, 1.0);
    }
}

void hc_decay2body(const double* e, const float* w32, uint64_t n_parents, double ma2, const alpk_event_t* ev,
                   uint32_t n_samples, const double* u, int fast, double* p1, double* p2, double* w) {
    const EventConst c = to_event(ev);
    const uint64_t n = n_parents * n_samples;
    for (uint64_t r = 0; r < n_parents; ++r) {
        const double ew = c.days_f64 ? (c.days_sec * (double)w32[r]) / (double)n_samples
                                     : (double)((c.days_sec32 * w32[r]) / (float)n_samples);
        double row[PR_NP];
        decay_parent_params(e[r], ma2, ew, row);
        for (uint32_t s = 0; s < n_samples; ++s) {
            const uint64_t i = r * n_samples + s;
            double q1[4], q2[4];
            if (fast) decay_event_fast(u[(r * 2) * n_samples + s], u[(r * 2 + 1) * n_samples + s], row, q1, q2);
            else decay_event(u[(r * 2) * n_samples + s], u[(r * 2 + 1) * n_samples + s], row, q1, q2);
            for (int k = 0; k < 4; ++k) { p1[k * n + i] = q1[k]; p2[k * n + i] = q2[k]; }
            w[i] = row[PR_W];
        }
    }
}

#include "hostcheck2.inc"

// Scatter2to2MC sample arithmetic (alp_math4.cuh); u = [2][n]: phi block then theta block
void hc_scatter2(const alpk_scatter2_t* h, uint64_t n, const double* u, double* t, double* w, double* p3cm, double* p3lab,
                 double* p4lab, double* cosine) {
    Scatter2Const c;
    c.m2.model = h->m2.model; c.m2.ma2 = h->m2.ma2; c.m2.ma4 = h->m2.ma4; c.m2.mN2 = h->m2.mN2; c.m2.mN4 = h->m2.mN4;
    c.m2.z = h->m2.z; c.m2.ff_a2 = h->m2.ff_a2; c.m2.pref = h->m2.pref; c.m2.me2 = h->m2.me2; c.m2.me4 = h->m2.me4;
    c.m2.tmax = h->m2.tmax;
    c.s = h->s; c.msum = h->msum; c.pp = h->pp; c.ee = h->ee; c.p3_cm = h->p3_cm; c.e3_cm = h->e3_cm; c.vol = h->vol;
    c.p1_cm = h->p1_cm; c.n_samples = h->n_samples; c.den = h->den; c.log_sampling = h->log_sampling;
    for (int k = 0; k < 16; ++k) c.mat[k] = h->mat[k];
    for (int k = 0; k < 4; ++k) c.p12[k] = h->p12[k];
    for (uint64_t i = 0; i < n; ++i) {
        Scatter2Out o;
        scatter2_sample(u[i], u[n + i], c, o);
        t[i] = o.t; w[i] = o.w; cosine[i] = o.cosine;
        for (int k = 0; k < 4; ++k) { p3cm[k * n + i] = o.p3cm[k]; p3lab[k * n + i] = o.p3lab[k]; p4lab[k * n + i] = o.p4lab[k]; }
    }
}

void hc_m2_eval(const alpk_m2_t* h, const double* s, const double* t, uint64_t n, double* out) {
    M2Const c;
    c.model = h->model; c.ma2 = h->ma2; c.ma4 = h->ma4; c.mN2 = h->mN2; c.mN4 = h->mN4; c.z = h->z; c.ff_a2 = h->ff_a2;
    c.pref = h->pref; c.me2 = h->me2; c.me4 = h->me4; c.tmax = h->tmax;
    for (uint64_t i = 0; i < n; ++i) out[i] = m2_eval(c, s[i], t[i]);
}

}  // extern "C"
