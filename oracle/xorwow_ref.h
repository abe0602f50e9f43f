/*
 * oracle/xorwow_ref.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Plain-C restatement of the XORWOW generator that the reference uses through
 * cuRAND's device API (`curandState` == curandStateXORWOW,
 * /usr/local/cuda/include/curand_kernel.h:150-156; cuRAND 10.3.10 / CUDA 12.9;
 * the dependency is NOT vendored under /root/reference).
 *
 *   seeding   : curand_kernel.h:800-825 (_curand_init_inplace)
 *   sequence  : curand_kernel.h:720-736 (_skipahead_sequence_inplace), one
 *               subsequence = 2^67 draws
 *   offset    : curand_kernel.h:700-718 (_skipahead_inplace)
 *   draw      : curand_kernel.h:863-874 (curand)
 *
 * The jump matrices are NOT copied from curand_precalc.h: they are rebuilt
 * from the published recurrence (Marsaglia xorshift, 5 words) by repeated
 * squaring over GF(2): T = one-step map, J = T^(2^67).
 * tests/test_oracle_rng.py pins this file against golden vectors produced by
 * cuRAND's own header (oracle/ref/gen_curand_golden.cu).
 */
#ifndef HW1F_ORACLE_XORWOW_REF_H
#define HW1F_ORACLE_XORWOW_REF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XW_WORDS 5
#define XW_BITS 160

/* A linear map on GF(2)^160: row b = image of basis vector e_b,
 * bit index b = 32*word + bit (same layout as cuRAND's __curand_matvec_inplace,
 * curand_kernel.h:316-334). */
typedef struct { uint32_t row[XW_BITS][XW_WORDS]; } xw_matrix;

typedef struct {
    uint32_t d;
    uint32_t v[XW_WORDS];
} xw_state;

/* Build (once) T^(2^k), k=0..XW_NPOW-1 and J^(2^k), k=0..XW_NPOW-1. Thread-safe
 * only if called once before use; xw_init() calls it lazily. */
#define XW_NPOW 48
void xw_build_tables(void);
const xw_matrix* xw_step_pow2(int k);  /* T^(2^k)          */
const xw_matrix* xw_seq_pow2(int k);   /* J^(2^k), J=T^(2^67) */

void xw_matvec(const xw_matrix* m, const uint32_t in[XW_WORDS], uint32_t out[XW_WORDS]);
void xw_matmul(const xw_matrix* a, const xw_matrix* b, xw_matrix* out); /* out = a o b */

/* curand_init(seed, subsequence, offset, &state) */
void xw_init(uint64_t seed, uint64_t subsequence, uint64_t offset, xw_state* st);
/* curand(&state) */
uint32_t xw_next(xw_state* st);

#ifdef __cplusplus
}
#endif
#endif
