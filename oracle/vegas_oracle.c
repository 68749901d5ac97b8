/*
 * ORACLE (test infrastructure, NOT product code) -- plain-C restatement of the
 * N3PDF/vegasflow v1.4.0 VEGAS hot path, fused per event, plus the
 * Philox4x32-10 stream definition shared with the CUDA kernels.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product never does.
 *
 * PARITY UNPINNED: the reference (Python on TensorFlow) cannot run in this
 * image and its tests carry no golden vectors for this path (only the
 * histogram scatter, tests/test_utils.py:11-30).  This file is pinned against
 * analytic answers and against oracle/vegas_ref.py (numpy, TF-shaped).
 *
 * Conventions (same as oracle/vegas_ref.py): reductions over the dimension
 * index run left to right; no FMA contraction on reference arithmetic
 * (build with -ffp-contract=off); float->int casts truncate.
 * Reference citations are file:line relative to /root/reference.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BINS_MAX 50 /* configflow.py:13 */
#define ALPHA 1.5   /* configflow.py:14 */
#define BETA 0.75   /* configflow.py:15 */
#define TECH_CUT 1e-8 /* configflow.py:16 */
#define VF_MAX_DIM 64

enum { VF_MODE_PLAIN = 0, VF_MODE_VEGAS = 1 };
enum { VF_INT_SYMGAUSS = 0, VF_INT_PRODUCT = 1 };

/* ------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  This is the      */
/* engine's own stream definition (the reference delegates to          */
/* tf.random.uniform, monte_carlo.py:264-266, whose stream cannot be   */
/* reproduced without TensorFlow).                                     */
/* ------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

void vfo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Two 32-bit words -> uniform double r in (TECH_CUT, 1-TECH_CUT]:
 *   m = 1.mantissa (52 bits: low 20 bits of `hi`, all of `lo`) in [1,2)
 *   v = fma(m, S, 3T)  in [1+T, 2-T)   with S = 1 - 2*TECH_CUT, T = TECH_CUT  (one rounding)
 *   r = 2 - v          exact: r is a multiple of 2^-52
 * so that 1 - r = v - 1 is exact as well and the fused CUDA kernel may evaluate the reference's
 * xn = 50*(1-r) (vflow.py:117) as fma(v, 50, -50) with a bit-identical result
 * (vegasflow_b200/csrc/vf_common.cuh::u52_to_v). */
static inline double u52_to_uniform(uint32_t hi, uint32_t lo) {
    union { uint64_t u; double d; } v;
    v.u = ((uint64_t)(0x3FF00000u | (hi & 0xFFFFFu)) << 32) | lo;
    const double S = 1.0 - 2.0 * TECH_CUT;
    const double C = 3.0 * TECH_CUT;
    return 2.0 - fma(v.d, S, C);
}

/* One word -> uniform in (TECH_CUT, 1-TECH_CUT): top 32 mantissa bits, m = 1 + k*2^-32,
 *   v = fma(m, S, 3T + S*2^-33) = 1 + T + (k + 1/2) * 2^-32 * S,  r = 2 - v  (32-bit stream) */
static inline double u32_to_uniform(uint32_t k) {
    union { uint64_t u; double d; } v;
    v.u = ((uint64_t)(0x3FF00000u | (k >> 12)) << 32) | (uint32_t)(k << 20);
    const double S = 1.0 - 2.0 * TECH_CUT;
    const double C = 3.0 * TECH_CUT + S * 1.1641532182693481e-10; /* 2^-33 */
    return 2.0 - fma(v.d, S, C);
}

/* Stream selection: 52 (default, two words per uniform) or 32 (one word per uniform). */
static int g_rng_bits = 52;
void vfo_set_rng_bits(int bits) { g_rng_bits = bits == 32 ? 32 : 52; }

/* Uniforms of event `ev` (global index), dims 2p and 2p+1 come from the
 * Philox block with counter (ev_lo, ev_hi, p, iteration), key = seed. */
static inline void event_uniforms(uint64_t seed, uint32_t iteration, uint64_t ev, int n_dim,
                                  double* r) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    if (g_rng_bits == 32) { /* block p feeds dims 4p .. 4p+3 */
        for (int p = 0; 4 * p < n_dim; ++p) {
            uint32_t ctr[4] = {(uint32_t)ev, (uint32_t)(ev >> 32), (uint32_t)p, iteration};
            uint32_t o[4];
            vfo_philox4x32_10(ctr, key, o);
            for (int h = 0; h < 4 && 4 * p + h < n_dim; ++h) r[4 * p + h] = u32_to_uniform(o[h]);
        }
        return;
    }
    for (int p = 0; 2 * p < n_dim; ++p) {
        uint32_t ctr[4] = {(uint32_t)ev, (uint32_t)(ev >> 32), (uint32_t)p, iteration};
        uint32_t o[4];
        vfo_philox4x32_10(ctr, key, o);
        r[2 * p] = u52_to_uniform(o[0], o[1]);
        if (2 * p + 1 < n_dim) r[2 * p + 1] = u52_to_uniform(o[2], o[3]);
    }
}

void vfo_uniforms(uint64_t seed, uint32_t iteration, uint64_t ev_begin, int64_t n, int n_dim,
                  double* rnds) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        event_uniforms(seed, iteration, ev_begin + (uint64_t)i, n_dim, rnds + i * n_dim);
}

/* ------------------------------------------------------------------ */
/* Integrands                                                          */
/* ------------------------------------------------------------------ */
typedef struct { double pref, C; } symgauss_consts;

/* examples/simgauss_tf.py:22-32 */
static symgauss_consts symgauss_setup(int n_dim) {
    symgauss_consts c;
    double a = 0.1;
    double n100 = 100.0 * n_dim;
    c.pref = pow(1.0 / a / sqrt(M_PI), (double)n_dim); /* :28 */
    c.C = (n100 + 1) * n100 / 2.0;                     /* :29 == :31 exactly */
    return c;
}

static inline double symgauss_eval(const double* x, int n_dim, symgauss_consts c) {
    const double a = 0.1;
    double s = 0.0;
    for (int j = 0; j < n_dim; ++j) {
        double t = (x[j] - 0.5) / a; /* :30 */
        double q = t * t;
        s = (j == 0) ? q : s + q;
    }
    double coef = c.C + s; /* :29-30 */
    coef = coef - c.C;     /* :31 */
    return c.pref * exp(-coef); /* :32 */
}

/* README.md:63-68 */
static inline double product_eval(const double* x, int n_dim) {
    double p = x[0];
    for (int j = 1; j < n_dim; ++j) p = p * x[j];
    return p;
}

void vfo_integrand(int integrand, int n_dim, int64_t n, const double* x, double* f) {
    symgauss_consts sc = symgauss_setup(n_dim);
    for (int64_t i = 0; i < n; ++i)
        f[i] = integrand == VF_INT_SYMGAUSS ? symgauss_eval(x + i * n_dim, n_dim, sc)
                                            : product_eval(x + i * n_dim, n_dim);
}

/* ------------------------------------------------------------------ */
/* a2+a3: vflow.py:93-126 + 39-83, one event                           */
/* ------------------------------------------------------------------ */
static inline double vegas_map_event(const double* r, const double* divisions, int n_dim, double* x,
                                     int32_t* ind) {
    double w = 1.0;
    for (int j = 0; j < n_dim; ++j) {
        double xn = 50.0 * (1.0 - r[j]);           /* vflow.py:117 */
        int32_t k = (int32_t)xn;                   /* :67 */
        const double* div = divisions + j * (BINS_MAX + 1);
        double x_ini = div[k];                     /* :70 */
        double x_fin = div[k + 1];                 /* :71 */
        double xdelta = x_fin - x_ini;             /* :73 */
        double aux = xn - floor(xn);               /* :75 */
        double prod = xdelta * aux;
        x[j] = x_ini + prod;                       /* :76 */
        double t = xdelta * 50.0;                  /* :78 */
        w = (j == 0) ? t : w * t;
        ind[j] = k;
    }
    return w;
}

/* a1: monte_carlo.py:270-274 */
static inline double apply_jac(double w, double xjac, int n_dim, double* x, const double* xmin,
                               const double* xdelta) {
    w = w * xjac;
    if (xdelta) {
        double jac = xdelta[0];
        for (int j = 1; j < n_dim; ++j) jac = jac * xdelta[j];
        for (int j = 0; j < n_dim; ++j) {
            double t = x[j] * xdelta[j];
            x[j] = xmin[j] + t;
        }
        w = w * jac;
    }
    return w;
}

/* Parity entry: externally supplied uniforms -> x, w, ind, w*f.
 * vflow.py:389-417 for mode VEGAS; plain.py:18-27 for mode PLAIN. */
void vfo_digest_from_uniforms(int mode, int integrand, int n_dim, int64_t n, const double* rnds,
                              const double* divisions, double xjac, const double* xmin,
                              const double* xdelta, double* x_out, double* w_out, int32_t* ind_out,
                              double* wf_out) {
    symgauss_consts sc = symgauss_setup(n_dim);
    for (int64_t i = 0; i < n; ++i) {
        double x[VF_MAX_DIM];
        int32_t ind[VF_MAX_DIM];
        double w;
        const double* r = rnds + i * n_dim;
        if (mode == VF_MODE_VEGAS) {
            w = vegas_map_event(r, divisions, n_dim, x, ind);
        } else {
            for (int j = 0; j < n_dim; ++j) { x[j] = r[j]; ind[j] = 0; }
            w = 1.0;
        }
        w = apply_jac(w, xjac, n_dim, x, xmin, xdelta);
        double f = integrand == VF_INT_SYMGAUSS ? symgauss_eval(x, n_dim, sc)
                                                : product_eval(x, n_dim);
        for (int j = 0; j < n_dim; ++j) {
            if (x_out) x_out[i * n_dim + j] = x[j];
            if (ind_out) ind_out[i * n_dim + j] = ind[j];
        }
        if (w_out) w_out[i] = w;
        if (wf_out) wf_out[i] = w * f;
    }
}

/* a4/a5/a6: fused event loop on the engine's Philox stream.
 * Accumulates (sum wf, sum (wf)^2) into out_sums[2] and, when train != 0 and
 * mode == VEGAS, sum (wf)^2 per (dim, bin) into out_hist[d*50]. */
void vfo_run_event(int mode, int integrand, int n_dim, uint64_t ev_begin, int64_t n_events,
                   double xjac, uint64_t seed, uint32_t iteration, int train,
                   const double* divisions, const double* xmin, const double* xdelta,
                   double* out_sums, double* out_hist, int nthreads) {
    symgauss_consts sc = symgauss_setup(n_dim);
    const int nh = n_dim * BINS_MAX;
    double tot = 0.0, tot2 = 0.0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel reduction(+ : tot, tot2)
    {
        double* hist = (double*)calloc((size_t)nh, sizeof(double));
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n_events; ++i) {
            double r[VF_MAX_DIM], x[VF_MAX_DIM];
            int32_t ind[VF_MAX_DIM];
            event_uniforms(seed, iteration, ev_begin + (uint64_t)i, n_dim, r);
            double w;
            if (mode == VF_MODE_VEGAS) {
                w = vegas_map_event(r, divisions, n_dim, x, ind);
            } else {
                for (int j = 0; j < n_dim; ++j) x[j] = r[j];
                w = 1.0;
            }
            w = apply_jac(w, xjac, n_dim, x, xmin, xdelta);
            double f = integrand == VF_INT_SYMGAUSS ? symgauss_eval(x, n_dim, sc)
                                                    : product_eval(x, n_dim);
            double tmp = w * f;      /* vflow.py:416 */
            double tmp2 = tmp * tmp; /* :417 */
            tot += tmp;              /* :420 */
            tot2 += tmp2;            /* :421 */
            if (train && mode == VF_MODE_VEGAS)
                for (int j = 0; j < n_dim; ++j) hist[j * BINS_MAX + ind[j]] += tmp2; /* utils.py:40-43 */
        }
        if (train && mode == VF_MODE_VEGAS) {
#pragma omp critical
            for (int k = 0; k < nh; ++k) out_hist[k] += hist[k];
        }
        free(hist);
    }
    out_sums[0] += tot;
    out_sums[1] += tot2;
}

/* ------------------------------------------------------------------ */
/* a8: refine_grid_per_dimension, vflow.py:135-211                     */
/* ------------------------------------------------------------------ */
static void refine_one(const double* t, double* sub) {
    double sm[BINS_MAX], wei[BINS_MAX], nb[BINS_MAX + 1];
    double sum_t = 0.0;
    for (int i = 0; i < BINS_MAX; ++i) {
        double left = i > 0 ? t[i - 1] : 0.0;              /* :156 */
        double right = i < BINS_MAX - 1 ? t[i + 1] : 0.0;
        double s = (t[i] + right) + left;                  /* :158 */
        double meaner = (i == 0 || i == BINS_MAX - 1) ? 2.0 : 3.0; /* :153-154 */
        sm[i] = fmax(s / meaner, 1e-30);                   /* :159 */
        sum_t = sum_t + sm[i];                             /* :162 */
    }
    double log_sum = log(sum_t);
    double ave = 0.0;
    for (int i = 0; i < BINS_MAX; ++i) {
        double aux = (1.0 - sm[i] / sum_t) / (log_sum - log(sm[i])); /* :163-164 */
        wei[i] = pow(aux, ALPHA);                                    /* :165 */
        ave = ave + wei[i];
    }
    ave = ave / BINS_MAX; /* :166 */
    nb[0] = 0.0;          /* :195 */
    double bw = 0.0, cur = 0.0, prev = 0.0;
    int n = -1;
    for (int k = 1; k < BINS_MAX; ++k) { /* :201 */
        while (bw < ave) {               /* :170-190 */
            n += 1;
            if (n > BINS_MAX - 1) { n = BINS_MAX - 1; break; } /* guard, SURVEY 8c */
            bw = bw + wei[n];
            prev = cur;
            cur = sub[n + 1];
        }
        bw = bw - ave;                          /* :205 */
        double delta = (cur - prev) * bw / wei[n]; /* :206 */
        nb[k] = cur - delta;                    /* :207 */
    }
    nb[BINS_MAX] = 1.0; /* :208 */
    memcpy(sub, nb, sizeof(nb));
}

/* VegasFlow.refine_grid, vflow.py:349-362 (in place) */
void vfo_refine_grid(int n_dim, const double* hist, double* divisions) {
    for (int j = 0; j < n_dim; ++j) refine_one(hist + j * BINS_MAX, divisions + j * (BINS_MAX + 1));
}

/* ------------------------------------------------------------------ */
/* VEGAS+: vflowplus.py:46-80 (a10), 187-220 (a11)                     */
/* Events are ordered by cube (vflowplus.py:67); ev_offset[c] is the   */
/* exclusive prefix sum of n_ev.  Cube c has lexicographic coordinates */
/* with dim 0 most significant (vflowplus.py:126-128).                 */
/* ------------------------------------------------------------------ */
void vfo_plus_run_event(int integrand, int n_dim, int n_strat, int64_t n_cubes,
                        const int32_t* n_ev, double xjac, uint64_t seed, uint32_t iteration,
                        int train, const double* divisions, const double* xmin,
                        const double* xdelta, const double* rnds_or_null, double* ress,
                        double* arr_var, double* out_hist, double* x_out, double* w_out,
                        int32_t* ind_out, double* wf_out) {
    symgauss_consts sc = symgauss_setup(n_dim);
    int64_t ev = 0;
    for (int64_t c = 0; c < n_cubes; ++c) {
        int coords[VF_MAX_DIM];
        int64_t rem = c;
        for (int j = n_dim - 1; j >= 0; --j) { coords[j] = (int)(rem % n_strat); rem /= n_strat; }
        double s1 = 0.0, s2 = 0.0;
        double fn = (double)n_ev[c];
        for (int32_t e = 0; e < n_ev[c]; ++e, ++ev) {
            double r[VF_MAX_DIM], x[VF_MAX_DIM];
            int32_t ind[VF_MAX_DIM];
            if (rnds_or_null) memcpy(r, rnds_or_null + ev * n_dim, sizeof(double) * n_dim);
            else event_uniforms(seed, iteration, (uint64_t)ev, n_dim, r);
            double w = 1.0;
            for (int j = 0; j < n_dim; ++j) {
                double xn = ((double)coords[j] + r[j]) * 50.0 / (double)n_strat; /* :72 */
                int32_t k = (int32_t)xn;
                const double* div = divisions + j * (BINS_MAX + 1);
                double x_ini = div[k], x_fin = div[k + 1];
                double xd = x_fin - x_ini;
                double aux = xn - floor(xn);
                double prod = xd * aux;
                x[j] = x_ini + prod;
                double t = xd * 50.0;
                w = (j == 0) ? t : w * t;
                ind[j] = k;
            }
            w = w / fn;                                   /* :77 */
            w = apply_jac(w, xjac, n_dim, x, xmin, xdelta); /* monte_carlo.py:270-274 */
            double f = integrand == VF_INT_SYMGAUSS ? symgauss_eval(x, n_dim, sc)
                                                    : product_eval(x, n_dim);
            double tmp = w * f;      /* vflowplus.py:209 */
            double tmp2 = tmp * tmp; /* :210 */
            s1 += tmp;               /* :213 */
            s2 += tmp2;              /* :214 */
            if (train)
                for (int j = 0; j < n_dim; ++j) out_hist[j * BINS_MAX + ind[j]] += tmp2;
            if (x_out) memcpy(x_out + ev * n_dim, x, sizeof(double) * n_dim);
            if (ind_out) memcpy(ind_out + ev * n_dim, ind, sizeof(int32_t) * n_dim);
            if (w_out) w_out[ev] = w;
            if (wf_out) wf_out[ev] = tmp;
        }
        ress[c] += s1;
        arr_var[c] = s2 * fn - s1 * s1; /* :216-217 (single chunk) */
    }
}
