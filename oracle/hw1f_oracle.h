/*
 * oracle/hw1f_oracle.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the reference's HW1F hot path (kernels + host
 * estimator algebra).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: the integer RNG layer is PINNED against cuRAND's own header
 * (tests/golden/xorwow_golden.json).  The floating-point layer follows the
 * operation order of the reference kernels as compiled with
 * `-O3 --use_fast_math -arch=sm_100` (read off the SASS), with the MUFU
 * approximations (rcp/lg2/ex2/sqrt/sin/cos) replaced by correctly rounded libm
 * calls; it is pinned against outputs of the reference binaries run on a B200
 * (tests/golden/ref_b200_*.json, produced by oracle/ref/capture_reference.py).
 */
#ifndef HW1F_ORACLE_H
#define HW1F_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors the compile-time configuration of include/common.cuh:16-39 and the
 * piecewise-linear theta of common.cuh:74-76,228-230. */
typedef struct {
    float a, sigma, r0;          /* H_A, H_SIGMA, H_R0                */
    float T_final;               /* T_FINAL                           */
    int n_steps;                 /* N_STEPS                           */
    int n_mat;                   /* N_MAT                             */
    float theta_a0, theta_b0;    /* theta(t)=a0+b0 t, t <  theta_break */
    float theta_a1, theta_b1;    /* theta(t)=a1+b1 t, t >= theta_break */
    float theta_break;
    float fd_theta_a1;           /* 0.014 quirk of src/3:385-387      */
} orc_params;

void orc_default_params(orc_params* p);

/* host-side model constants ------------------------------------------------ */
float orc_dt(const orc_params* p);                 /* common.cuh:33  */
float orc_mat_spacing(const orc_params* p);        /* common.cuh:34  */
float orc_exp_adt(const orc_params* p);            /* common.cuh:93  */
float orc_sig_st(const orc_params* p, float sigma);/* common.cuh:87-89 */
void orc_drift_tables(const orc_params* p, float sigma, float* drift, float* sigma_drift); /* common.cuh:60-84 */
void orc_shifted_drift_table(const orc_params* p, float sigma_new, float sigma_old, float* drift); /* src/3:374-398 */
int orc_steps_to(const orc_params* p, float S1);   /* (int)(S1/d_dt), common.cuh:322 with an exact reciprocal */

/* RNG ---------------------------------------------------------------------- */
void orc_draws(uint64_t seed, uint64_t subsequence, uint64_t offset, int n, uint32_t* out);
/* curand_normal stream (curand_normal.h:313-326) starting at normal index `offset_normals` */
void orc_normals(uint64_t seed, uint64_t subsequence, uint64_t offset_normals, int n, float* out);

/* Q1: simulate_zcb (market_data.cuh:25-79).  sum[m], m=0..n_mat-1, double sums of the
 * per-thread float p0_m; sumsq[m] = sum of p0_m^2 (may be NULL).  Entry 0 is left 0
 * (the reference overwrites it with 2N in the kernel; see orc_curve_finalize). */
void orc_bond_curve_sums(const orc_params* p, float sig_st, const float* drift,
                         uint64_t seed, uint64_t first_path, int64_t n_pairs,
                         uint64_t offset_normals, double* sum, double* sumsq);
/* compute_average_and_forward (market_data.cuh:101-127); P_sum_f[0] is replaced by 2*n_pairs */
void orc_curve_finalize(const orc_params* p, const float* P_sum_f, int64_t n_pairs, float* P, float* f);
/* recover_theta (src/2:14-35) */
void orc_theta(const orc_params* p, float sigma, const float* f, float* theta_rec, float* theta_orig, float* Ts);

/* Q2b: simulate_ZBC_control_variate (common.cuh:286-409): mom = {SX, SY, SXX, SYY, SXY} */
void orc_zbc_moments(const orc_params* p, float sigma, float sig_st, const float* drift,
                     uint64_t seed, uint64_t first_path, int64_t n_pairs, uint64_t offset_normals,
                     int n_steps_S1, float S1, float S2, float K,
                     const float* P_mkt, const float* f_mkt, double mom[5]);

typedef struct {
    float mean_X, mean_Y, var_Y, var_X, cov, beta, price_cv, corr_single, corr, control_adjustment;
} orc_zbc_result;
/* host algebra of src/2:154-179 and src/2:259-290 in float32 */
void orc_zbc_algebra(const float mom_f[5], int n_total, float P0S2, orc_zbc_result* out);

/* Q3: simulate_sensitivity (src/3:22-96): sum of (term1-term2), and sum of squares */
void orc_vega_pathwise_sums(const orc_params* p, float sigma, float sig_st,
                            const float* drift, const float* sigma_drift,
                            uint64_t seed, uint64_t first_path, int64_t n_paths, uint64_t offset_normals,
                            int n_steps_S1, float S1, float S2, float K,
                            const float* P_mkt, const float* f_mkt, double* sum, double* sumsq);

/* simulate_paths_show (market_data.cuh:136-160): out[n_show*(n_steps+1)] */
void orc_sample_paths(const orc_params* p, float sig_st, const float* drift,
                      uint64_t seed, uint64_t first_path, int n_show, uint64_t offset_normals, float* out);

/* 20-run statistics (src/2:305-324, src/3:570-589): out = {mean, var, sd, se, moe, lo, hi, cv_pct} */
void orc_run_stats(const float* samples, int n, float out[8]);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
