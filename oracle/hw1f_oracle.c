/*
 * oracle/hw1f_oracle.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED.  The integer layer (xorwow_ref.c) is bit-exact against cuRAND's own header executed on the
 * host (tests/golden/xorwow_golden.json, oracle/ref/gen_curand_golden.cu); the float layer is checked against the
 * outputs of the unmodified reference kernels run on a B200 (tests/golden/ref_b200_seed20251018.json, produced by
 * oracle/_ref/ref_harness; tests/test_oracle_vs_reference_fixture.py) and against closed-form Hull-White values.
 *
 * Build with -ffp-contract=off: every fused multiply-add below is an explicit
 * fmaf() placed where the reference's sm_100 SASS has an FFMA; everything else
 * rounds once per operation like the FMUL/FADD it restates.  FTZ/DAZ is enabled
 * around every entry point because the reference is built with --use_fast_math
 * (makefile:2).
 */
#include "hw1f_oracle.h"
#include "xorwow_ref.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__x86_64__) || defined(__i386__)
#include <xmmintrin.h>
#include <pmmintrin.h>
#define FTZ_SCOPE_BEGIN unsigned int _csr = _mm_getcsr(); _mm_setcsr(_csr | 0x8040u);
#define FTZ_SCOPE_END _mm_setcsr(_csr);
#define FTZ_THREAD _mm_setcsr(_mm_getcsr() | 0x8040u);
#else
#define FTZ_SCOPE_BEGIN
#define FTZ_SCOPE_END
#define FTZ_THREAD
#endif

/* ---- stand-ins for the MUFU approximations used by the fast-math build ----- */
#define LOG2E_F 1.4426950216293334961f   /* FMUL constant seen in SASS */
#define LN2_F 0.69314718246459960938f
#define INV_2PI_F 0.15915493667125701904f

static inline float mufu_rcp(float x) { return 1.0f / x; }
static inline float mufu_ex2(float x) { return exp2f(x); }
static inline float mufu_lg2(float x) { return log2f(x); }
static inline float mufu_sqrt(float x) { return sqrtf(x); }
/* MUFU.SIN/COS take the angle in revolutions (the compiler emits FMUL.RZ by 1/2pi first) */
static inline float mufu_sin_rev(float w) { return (float)sin(6.283185307179586476925 * (double)w); }
static inline float mufu_cos_rev(float w) { return (float)cos(6.283185307179586476925 * (double)w); }
static inline float fmul_rz(float a, float b)
{
    double p = (double)a * (double)b; /* exact: 24+24 bits */
    float f = (float)p;
    if (fabs((double)f) > fabs(p)) f = nextafterf(f, 0.0f);
    return f;
}
static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float exp_fast(float x) { return mufu_ex2(x * LOG2E_F); }   /* expf under --use_fast_math */

/* ---- parameters ------------------------------------------------------------ */
void orc_default_params(orc_params* p)
{
    p->a = 1.0f; p->sigma = 0.1f; p->r0 = 0.012f;       /* common.cuh:37-39 */
    p->T_final = 10.0f; p->n_steps = 1000; p->n_mat = 101; /* common.cuh:16-22 */
    p->theta_a0 = 0.012f; p->theta_b0 = 0.0014f;          /* common.cuh:229  */
    p->theta_a1 = 0.019f; p->theta_b1 = 0.001f;
    p->theta_break = 5.0f;
    p->fd_theta_a1 = 0.014f;                              /* src/3:387       */
}

float orc_dt(const orc_params* p) { return p->T_final / p->n_steps; }
float orc_mat_spacing(const orc_params* p) { return p->T_final / (p->n_mat - 1); }
float orc_exp_adt(const orc_params* p) { return expf(-p->a * orc_dt(p)); }
float orc_sig_st(const orc_params* p, float sigma)
{
    return sigma * sqrtf((1.0f - expf(-2.0f * p->a * orc_dt(p))) / (2.0f * p->a));
}

void orc_drift_tables(const orc_params* p, float sigma, float* drift, float* sigma_drift)
{
    const float H_A = p->a, H_DT = orc_dt(p);
    float h_exp_adt = expf(-H_A * H_DT);
    float om_a = (1.0f - h_exp_adt) / H_A;
    float om_a_sq = om_a / H_A;
    for (int i = 0; i < p->n_steps; i++) {
        float s = i * H_DT;
        float t = (i + 1) * H_DT;
        float first_term = ((s + H_DT) - h_exp_adt * s) / H_A - om_a_sq;
        drift[i] = (s < p->theta_break) ? (p->theta_b0 * first_term + p->theta_a0 * om_a)
                                        : (p->theta_b1 * first_term + p->theta_a1 * om_a);
        if (sigma_drift) {
            float sigma_term = (2.0f * sigma * expf(-H_A * t)) * (coshf(H_A * t) - coshf(H_A * s));
            sigma_drift[i] = sigma_term / (H_A * H_A);
        }
    }
}

void orc_shifted_drift_table(const orc_params* p, float sigma_new, float sigma_old, float* out)
{
    const float H_A = p->a, H_DT = orc_dt(p);
    float shift_coeff = (sigma_new * sigma_new - sigma_old * sigma_old) / (2.0f * H_A);
    float h_exp_adt = expf(-H_A * H_DT);
    float om_a = (1.0f - h_exp_adt) / H_A;
    float om_a_sq = om_a / H_A;
    for (int i = 0; i < p->n_steps; i++) {
        float s = i * H_DT;
        float t = (i + 1) * H_DT;
        float first_term = ((s + H_DT) - h_exp_adt * s) / H_A - om_a_sq;
        float base = (s < p->theta_break) ? (p->theta_b0 * first_term + p->theta_a0 * om_a)
                                          : (p->theta_b1 * first_term + p->fd_theta_a1 * om_a);
        float adj = (shift_coeff / H_A) *
                    (1.0f + expf(-2.0f * H_A * t) - expf(-H_A * (t - s)) - expf(-H_A * (t + s)));
        out[i] = base + adj;
    }
}

int orc_steps_to(const orc_params* p, float S1)
{
    FTZ_SCOPE_BEGIN
    int n = (int)(mufu_rcp(orc_dt(p)) * S1);
    FTZ_SCOPE_END
    return n;
}

/* ---- RNG --------------------------------------------------------------------- */
void orc_draws(uint64_t seed, uint64_t subsequence, uint64_t offset, int n, uint32_t* out)
{
    xw_state st;
    xw_init(seed, subsequence, offset, &st);
    for (int i = 0; i < n; i++) out[i] = xw_next(&st);
}

/* _curand_box_muller (curand_normal.h:70-87) as compiled for the device */
static inline void box_muller(uint32_t x, uint32_t y, float* n_sin, float* n_cos)
{
    float u = fmaf((float)x, as_float(0x2f800000u), as_float(0x2f000000u));
    float v = fmaf((float)y, as_float(0x30c90fdbu), as_float(0x30490fdbu));
    float s = mufu_sqrt((mufu_lg2(u) * LN2_F) * -2.0f);
    float w = fmul_rz(v, INV_2PI_F);
    *n_sin = s * mufu_sin_rev(w);
    *n_cos = s * mufu_cos_rev(w);
}

typedef struct {
    xw_state st;
    int have_extra;
    float extra;
} normal_stream;

static void ns_init(normal_stream* ns, uint64_t seed, uint64_t subsequence, uint64_t offset_normals)
{
    xw_init(seed, subsequence, 2 * (offset_normals / 2), &ns->st);
    ns->have_extra = 0;
    ns->extra = 0.0f;
    if (offset_normals & 1) { /* the cached cos-branch value is next (curand_normal.h:324-325) */
        float a, b;
        uint32_t x = xw_next(&ns->st), y = xw_next(&ns->st);
        box_muller(x, y, &a, &b);
        ns->have_extra = 1;
        ns->extra = b;
    }
}

static inline float ns_next(normal_stream* ns)
{
    if (!ns->have_extra) {
        float a, b;
        uint32_t x = xw_next(&ns->st), y = xw_next(&ns->st);
        box_muller(x, y, &a, &b);
        ns->extra = b;
        ns->have_extra = 1;
        return a;
    }
    ns->have_extra = 0;
    return ns->extra;
}

void orc_normals(uint64_t seed, uint64_t subsequence, uint64_t offset_normals, int n, float* out)
{
    FTZ_SCOPE_BEGIN
    normal_stream ns;
    ns_init(&ns, seed, subsequence, offset_normals);
    for (int i = 0; i < n; i++) out[i] = ns_next(&ns);
    FTZ_SCOPE_END
}

/* evolve_hull_white_step (common.cuh:237-244) as compiled: FFMA, FADD, FMUL, FFMA */
static inline void hw_step(float* r, float* integral, float shock, float exp_adt, float dt)
{
    float r_next = fmaf(*r, exp_adt, shock);
    float h = (r_next + *r) * 0.5f;
    *integral = fmaf(h, dt, *integral);
    *r = r_next;
}

/* ---- Q1 ------------------------------------------------------------------------ */
void orc_bond_curve_sums(const orc_params* p, float sig_st, const float* drift,
                         uint64_t seed, uint64_t first_path, int64_t n_pairs,
                         uint64_t offset_normals, double* sum, double* sumsq)
{
    xw_build_tables();
    FTZ_SCOPE_BEGIN
    const int n_mat = p->n_mat, n_steps = p->n_steps;
    const int stride = n_steps / (n_mat - 1); /* SAVE_STRIDE, common.cuh:29 */
    const float exp_adt = orc_exp_adt(p), dt = orc_dt(p), r0 = p->r0;
    for (int m = 0; m < n_mat; m++) { sum[m] = 0.0; if (sumsq) sumsq[m] = 0.0; }
#pragma omp parallel
    {
        FTZ_THREAD
        double* ls = (double*)calloc(2 * (size_t)n_mat, sizeof(double));
        double* lq = ls + n_mat;
#pragma omp for schedule(static)
        for (int64_t q = 0; q < n_pairs; q++) {
            normal_stream ns;
            ns_init(&ns, seed, first_path + (uint64_t)q, offset_normals);
            float r1 = r0, r2 = r0, I1 = 0.0f, I2 = 0.0f;
            for (int i = 1; i <= n_steps; i++) {
                float d = drift[i - 1];
                float G = ns_next(&ns);
                float sp = fmaf(G, sig_st, d);
                float sm = fmaf(-G, sig_st, d);
                hw_step(&r1, &I1, sp, exp_adt, dt);
                hw_step(&r2, &I2, sm, exp_adt, dt);
                if (i % stride == 0) {
                    int m = i / stride;
                    if (m < n_mat) {
                        float p0 = mufu_ex2(I1 * -LOG2E_F) + mufu_ex2(I2 * -LOG2E_F);
                        ls[m] += (double)p0;
                        lq[m] += (double)p0 * (double)p0;
                    }
                }
            }
        }
#pragma omp critical
        {
            for (int m = 0; m < n_mat; m++) { sum[m] += ls[m]; if (sumsq) sumsq[m] += lq[m]; }
        }
        free(ls);
    }
    FTZ_SCOPE_END
}

void orc_curve_finalize(const orc_params* p, const float* P_sum_f, int64_t n_pairs, float* P, float* f)
{
    FTZ_SCOPE_BEGIN
    const int n_mat = p->n_mat;
    const int n_paths = (int)(2 * n_pairs);
    const float inv_dT = 1.0f / orc_mat_spacing(p); /* host division, src/1:76 */
    const float rn = mufu_rcp((float)n_paths);
    for (int m = 0; m < n_mat; m++) {
        float s = (m == 0) ? 2.0f * (float)n_pairs : P_sum_f[m]; /* market_data.cuh:76-78 */
        P[m] = s * rn;
    }
    for (int m = 0; m < n_mat; m++) {
        int first = (m == 0) ? 0 : m - 1;
        int last = (m == n_mat - 1) ? n_mat - 1 : m + 1;
        float nscale = ((m == 0) || (m == n_mat - 1)) ? -1.0f : -0.5f;
        float c = nscale * inv_dT;
        float lf = mufu_lg2(P[first]) * LN2_F;
        float dl = fmaf(mufu_lg2(P[last]), LN2_F, -lf);
        f[m] = c * dl;
    }
    FTZ_SCOPE_END
}

void orc_theta(const orc_params* p, float sigma, const float* f, float* theta_rec, float* theta_orig, float* Ts)
{
    FTZ_SCOPE_BEGIN
    const int n = p->n_mat;
    const float a = p->a, sp = orc_mat_spacing(p);
    const float coef = (sigma * sigma) * mufu_rcp(a + a);
    const float m2a = a * -2.0f;
    for (int i = 0; i < n; i++) {
        float T = (float)i * sp;
        float df;
        if (i == 0) df = (f[1] - f[0]) * mufu_rcp(sp);
        else if (i == n - 1) df = (f[i] - f[i - 1]) * mufu_rcp(sp);
        else df = (f[i + 1] - f[i - 1]) * mufu_rcp(sp + sp);
        float e = mufu_ex2((m2a * T) * LOG2E_F);
        float om = 1.0f - e;
        theta_rec[i] = fmaf(coef, om, fmaf(f[i], a, df));
        theta_orig[i] = (T < p->theta_break) ? fmaf(T, p->theta_b0, p->theta_a0)
                                             : fmaf(T, p->theta_b1, p->theta_a1);
        Ts[i] = T;
    }
    FTZ_SCOPE_END
}

/* ---- bond formula pieces shared by Q2b/Q3 (common.cuh:180-225) ------------------- */
static float interp_mkt(const float* data, float T, int n_mat, float inv_spacing, float neg_spacing)
{
    int idx = (int)(T * inv_spacing);                /* T / spacing folded to T * 10 */
    if (idx >= n_mat - 1) return data[n_mat - 1];
    float al = fmaf((float)idx, neg_spacing, T) * inv_spacing;
    float om = 1.0f - al;
    return fmaf(data[idx], om, al * data[idx + 1]);
}

typedef struct { float B, A, om2, negB; } bond_consts;

static bond_consts bond_setup(const orc_params* p, float sigma, float S1, float S2,
                              const float* P_mkt, const float* f_mkt)
{
    bond_consts c;
    const float a = p->a;
    const float sp = orc_mat_spacing(p);
    const float inv_sp = 1.0f / sp;  /* compile-time fold in the reference: x/0.1f -> x*10.0f */
    float B = (1.0f - mufu_ex2(((S2 - S1) * a) * -LOG2E_F)) * mufu_rcp(a);
    float P0T = interp_mkt(P_mkt, S2, p->n_mat, inv_sp, -sp);
    float P0t = interp_mkt(P_mkt, S1, p->n_mat, inv_sp, -sp);
    float f0t = interp_mkt(f_mkt, S1, p->n_mat, inv_sp, -sp);
    float om2 = 1.0f - mufu_ex2(((a * -2.0f) * S1) * LOG2E_F);
    float t3 = ((sigma * sigma) * mufu_rcp(a * 4.0f)) * om2;
    t3 = t3 * B;
    t3 = t3 * B;
    float E = mufu_ex2(fmaf(f0t, B, -t3) * LOG2E_F);
    float ratio = mufu_rcp(P0t) * P0T;
    c.B = B; c.negB = -B; c.A = ratio * E; c.om2 = om2;
    return c;
}

/* ---- Q2b -------------------------------------------------------------------------- */
void orc_zbc_moments(const orc_params* p, float sigma, float sig_st, const float* drift,
                     uint64_t seed, uint64_t first_path, int64_t n_pairs, uint64_t offset_normals,
                     int n_steps_S1, float S1, float S2, float K,
                     const float* P_mkt, const float* f_mkt, double mom[5])
{
    xw_build_tables();
    FTZ_SCOPE_BEGIN
    const bond_consts bc = bond_setup(p, sigma, S1, S2, P_mkt, f_mkt);
    const float exp_adt = orc_exp_adt(p), dt = orc_dt(p), r0 = p->r0;
    double m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
#pragma omp parallel reduction(+ : m0, m1, m2, m3, m4)
    {
        FTZ_THREAD
#pragma omp for schedule(static)
        for (int64_t q = 0; q < n_pairs; q++) {
            normal_stream ns;
            ns_init(&ns, seed, first_path + (uint64_t)q, offset_normals);
            float r1 = r0, r2 = r0, I1 = 0.0f, I2 = 0.0f;
            for (int i = 0; i < n_steps_S1; i++) {
                float d = drift[i];
                float G = ns_next(&ns);
                hw_step(&r1, &I1, fmaf(G, sig_st, d), exp_adt, dt);
                hw_step(&r2, &I2, fmaf(-G, sig_st, d), exp_adt, dt);
            }
            float P1 = bc.A * mufu_ex2((r1 * bc.negB) * LOG2E_F);
            float P2 = bc.A * mufu_ex2((r2 * bc.negB) * LOG2E_F);
            float d1 = mufu_ex2(I1 * -LOG2E_F);
            float d2 = mufu_ex2(I2 * -LOG2E_F);
            float c1 = P1 * d1, c2 = P2 * d2;
            float x1 = d1 * fmaxf(0.0f, P1 - K);
            float x2 = d2 * fmaxf(0.0f, P2 - K);
            float tX = x1 + x2;
            float tY = c1 + c2;
            float tXX = fmaf(x1, x1, x2 * x2);
            float tYY = fmaf(c1, c1, c2 * c2);
            float tXY = fmaf(c1, x1, c2 * x2);
            m0 += tX; m1 += tY; m2 += tXX; m3 += tYY; m4 += tXY;
        }
    }
    mom[0] = m0; mom[1] = m1; mom[2] = m2; mom[3] = m3; mom[4] = m4;
    FTZ_SCOPE_END
}

void orc_zbc_algebra(const float mom_f[5], int n_total, float P0S2, orc_zbc_result* out)
{
    /* src/2:154-179 (single run) and src/2:259-290 (validation run); plain IEEE
     * float32 host arithmetic, no contraction (host code is built without -mfma). */
    float h_ZBC = mom_f[0], h_control = mom_f[1], h_ZBC_sq = mom_f[2], h_control_sq = mom_f[3], h_cross = mom_f[4];
    float mean_ZBC = h_ZBC / n_total;
    float mean_control = h_control / n_total;
    float E_Y2 = h_control_sq / n_total;
    float E_Y_sq = mean_control * mean_control;
    float var_control = E_Y2 - E_Y_sq;
    float E_XY = h_cross / n_total;
    float E_X_E_Y = mean_ZBC * mean_control;
    float cov = E_XY - E_X_E_Y;
    float beta = cov / var_control;
    float control_adjustment = beta * (mean_control - P0S2);
    float adjusted = mean_ZBC - control_adjustment;
    float corr_single = cov / (sqrtf(var_control) * sqrtf(E_Y2 - E_Y_sq)); /* src/2:178 (== beta) */
    float E_X2 = h_ZBC_sq / n_total;
    float var_ZBC = E_X2 - mean_ZBC * mean_ZBC;
    float corr = cov / sqrtf(var_ZBC * var_control);                        /* src/2:281 */
    out->mean_X = mean_ZBC; out->mean_Y = mean_control; out->var_Y = var_control; out->var_X = var_ZBC;
    out->cov = cov; out->beta = beta; out->price_cv = adjusted; out->corr_single = corr_single;
    out->corr = corr; out->control_adjustment = control_adjustment;
}

/* ---- Q3 pathwise --------------------------------------------------------------------- */
void orc_vega_pathwise_sums(const orc_params* p, float sigma, float sig_st,
                            const float* drift, const float* sigma_drift,
                            uint64_t seed, uint64_t first_path, int64_t n_paths, uint64_t offset_normals,
                            int n_steps_S1, float S1, float S2, float K,
                            const float* P_mkt, const float* f_mkt, double* sum, double* sumsq)
{
    xw_build_tables();
    FTZ_SCOPE_BEGIN
    const bond_consts bc = bond_setup(p, sigma, S1, S2, P_mkt, f_mkt);
    const float exp_adt = orc_exp_adt(p), dt = orc_dt(p), r0 = p->r0, a = p->a;
    const float c_t = mufu_rcp(sigma) * sig_st;              /* (d_sig_st / d_sigma) */
    const float xk = bc.om2 * (mufu_rcp(a + a) * sigma);     /* sigma/(2a) * (1-e^{-2aS1}) */
    double s = 0, sq = 0;
#pragma omp parallel reduction(+ : s, sq)
    {
        FTZ_THREAD
#pragma omp for schedule(static)
        for (int64_t q = 0; q < n_paths; q++) {
            normal_stream ns;
            ns_init(&ns, seed, first_path + (uint64_t)q, offset_normals);
            float r = r0, t = 0.0f, Ir = 0.0f, It = 0.0f;
            for (int i = 1; i <= n_steps_S1; i++) {
                float G = ns_next(&ns);
                float sr = fmaf(G, sig_st, drift[i - 1]);
                float stt = fmaf(c_t, G, sigma_drift[i - 1]);
                hw_step(&r, &Ir, sr, exp_adt, dt);
                hw_step(&t, &It, stt, exp_adt, dt);
            }
            float P = bc.A * mufu_ex2((r * bc.negB) * LOG2E_F);
            float disc = mufu_ex2(Ir * -LOG2E_F);
            float term1 = 0.0f;
            if (P > K) {
                float inner = fmaf(xk, bc.B, t);
                float y = (P * bc.negB) * inner;
                term1 = disc * y;
            }
            float payoff = fmaxf(0.0f, P - K);
            float z = disc * It;
            float v = fmaf(payoff, -z, term1);
            s += v; sq += (double)v * (double)v;
        }
    }
    *sum = s;
    if (sumsq) *sumsq = sq;
    FTZ_SCOPE_END
}

void orc_sample_paths(const orc_params* p, float sig_st, const float* drift,
                      uint64_t seed, uint64_t first_path, int n_show, uint64_t offset_normals, float* out)
{
    xw_build_tables();
    FTZ_SCOPE_BEGIN
    const float exp_adt = orc_exp_adt(p), dt = orc_dt(p);
    const int n_steps = p->n_steps;
    for (int q = 0; q < n_show; q++) {
        normal_stream ns;
        ns_init(&ns, seed, first_path + (uint64_t)q, offset_normals);
        float r = p->r0, I = 0.0f;
        out[(size_t)q * (n_steps + 1)] = r;
        for (int i = 1; i <= n_steps; i++) {
            float G = ns_next(&ns);
            hw_step(&r, &I, fmaf(G, sig_st, drift[i - 1]), exp_adt, dt);
            out[(size_t)q * (n_steps + 1) + i] = r;
        }
    }
    FTZ_SCOPE_END
}

void orc_run_stats(const float* samples, int n, float out[8])
{
    float mean = 0.0f;
    for (int i = 0; i < n; i++) mean += samples[i];
    mean /= n;
    float variance = 0.0f;
    for (int i = 0; i < n; i++) { float d = samples[i] - mean; variance += d * d; }
    variance /= (n - 1);
    float sd = sqrtf(variance);
    float se = sd / sqrtf((float)n);
    float moe = 2.093f * se;          /* t_{0.975,19}, src/2:320 */
    out[0] = mean; out[1] = variance; out[2] = sd; out[3] = se; out[4] = moe;
    out[5] = mean - moe; out[6] = mean + moe; out[7] = 100.0f * sd / mean;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
