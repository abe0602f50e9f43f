/*
 * oracle/xorwow_ref.c -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 * See xorwow_ref.h for the reference citations.
 */
#include "xorwow_ref.h"
#include <string.h>

static xw_matrix g_step[XW_NPOW]; /* T^(2^k) */
static xw_matrix g_seq[XW_NPOW];  /* J^(2^k) */
static int g_built = 0;

/* the xorshift part of curand() (curand_kernel.h:863-874), on v only */
static void step_v(uint32_t v[XW_WORDS])
{
    uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1];
    v[1] = v[2];
    v[2] = v[3];
    v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}

void xw_matvec(const xw_matrix* m, const uint32_t in[XW_WORDS], uint32_t out[XW_WORDS])
{
    uint32_t acc[XW_WORDS] = {0, 0, 0, 0, 0};
    for (int i = 0; i < XW_WORDS; i++) {
        uint32_t w = in[i];
        for (int j = 0; j < 32; j++) {
            if (w & (1u << j)) {
                const uint32_t* r = m->row[32 * i + j];
                for (int k = 0; k < XW_WORDS; k++) acc[k] ^= r[k];
            }
        }
    }
    memcpy(out, acc, sizeof(acc));
}

void xw_matmul(const xw_matrix* a, const xw_matrix* b, xw_matrix* out)
{
    /* (a o b) e_r = a (b e_r) */
    xw_matrix tmp;
    for (int r = 0; r < XW_BITS; r++) xw_matvec(a, b->row[r], tmp.row[r]);
    memcpy(out, &tmp, sizeof(tmp));
}

void xw_build_tables(void)
{
    if (g_built) return;
    /* T: apply one step to each basis vector */
    for (int b = 0; b < XW_BITS; b++) {
        uint32_t e[XW_WORDS] = {0, 0, 0, 0, 0};
        e[b >> 5] = 1u << (b & 31);
        step_v(e);
        memcpy(g_step[0].row[b], e, sizeof(e));
    }
    for (int k = 1; k < XW_NPOW; k++) xw_matmul(&g_step[k - 1], &g_step[k - 1], &g_step[k]);
    /* J = T^(2^67): continue squaring from T^(2^(XW_NPOW-1)) */
    xw_matrix m;
    memcpy(&m, &g_step[XW_NPOW - 1], sizeof(m));
    for (int k = XW_NPOW - 1; k < 67; k++) xw_matmul(&m, &m, &m);
    memcpy(&g_seq[0], &m, sizeof(m));
    for (int k = 1; k < XW_NPOW; k++) xw_matmul(&g_seq[k - 1], &g_seq[k - 1], &g_seq[k]);
    g_built = 1;
}

const xw_matrix* xw_step_pow2(int k) { xw_build_tables(); return &g_step[k]; }
const xw_matrix* xw_seq_pow2(int k) { xw_build_tables(); return &g_seq[k]; }

void xw_init(uint64_t seed, uint64_t subsequence, uint64_t offset, xw_state* st)
{
    xw_build_tables();
    /* curand_kernel.h:805-819 */
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    st->d = 6615241u + t1 + t0;
    st->v[0] = 123456789u + t0;
    st->v[1] = 362436069u ^ t0;
    st->v[2] = 521288629u + t1;
    st->v[3] = 88675123u ^ t1;
    st->v[4] = 5783321u + t0;
    /* subsequence jump: J^subsequence (curand applies base-4 digits of powers
     * J^(4^k); all powers of T commute so binary digits give the same map) */
    for (int k = 0; subsequence != 0 && k < XW_NPOW; k++, subsequence >>= 1)
        if (subsequence & 1) xw_matvec(&g_seq[k], st->v, st->v);
    /* offset jump: T^offset, d += 362437*offset (curand_kernel.h:700-718) */
    st->d += 362437u * (uint32_t)offset;
    for (int k = 0; offset != 0 && k < XW_NPOW; k++, offset >>= 1)
        if (offset & 1) xw_matvec(&g_step[k], st->v, st->v);
}

uint32_t xw_next(xw_state* st)
{
    step_v(st->v);
    st->d += 362437u;
    return st->v[4] + st->d;
}
