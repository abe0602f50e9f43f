/*
 * oracle/ref/gen_curand_golden.cu -- TEST INFRASTRUCTURE.
 *
 * Generates tests/golden/xorwow_golden.json by executing cuRAND's OWN header
 * (/usr/local/cuda/include/curand_kernel.h, cuRAND 10.3.10 / CUDA 12.9) on the
 * HOST: QUALIFIERS is redefined so that curand_init()/curand() compile as
 * host functions (the header carries a host copy of the jump tables,
 * precalc_xorwow_matrix_host).  This is the third-party generator the
 * reference calls at include/common.cuh:277-280 (curand_init(seed, idx, 0))
 * and common.cuh:327 / market_data.cuh:45 (curand_normal -> curand()).
 *
 * Build+run (no GPU needed):  make -C oracle golden
 */
#define QUALIFIERS static inline __host__ __device__
#include <curand_kernel.h>
#include <cstdio>
#include <cstdint>

struct Case { unsigned long long seed, seq, off; };

int main(int argc, char** argv)
{
    const Case cases[] = {
        {0ull, 0ull, 0ull},
        {1234ull, 0ull, 0ull},
        {1234ull, 1ull, 0ull},
        {1234ull, 5ull, 0ull},
        {1234ull, 1023ull, 0ull},
        {1234ull, 1024ull, 0ull},
        {1234ull, 65535ull, 0ull},
        {1234ull, 1048575ull, 0ull},
        {1234ull, 1048576ull, 0ull},
        {1234ull, (1ull << 30) - 1ull, 0ull},
        {1234ull, (1ull << 31) + 12345ull, 0ull},
        {1700000000ull, 0ull, 0ull},
        {1700000000ull, 1048575ull, 0ull},
        {1700000000037035ull, 1ull, 0ull},      /* hi32 of the seed != 0 */
        {0xffffffffffffffffull, 77ull, 0ull},
        {1234ull, 5ull, 500ull},                /* Q3 FD window   (src/3:407-435) */
        {1234ull, 5ull, 1000ull},               /* Q3 recalibrated window         */
        {1234ull, 1048575ull, 2000ull},
        {42ull, 31ull, 499ull},
        {42ull, 777ull, 1000003ull},
    };
    const int n_cases = (int)(sizeof(cases) / sizeof(cases[0]));
    const int n_head = 8;
    FILE* f = (argc > 1) ? fopen(argv[1], "w") : stdout;
    if (!f) return 1;
    fprintf(f, "{\n  \"generator\": \"cuRAND host path of curand_kernel.h (CURAND_VERSION %d)\",\n",
            (int)CURAND_VERSION);
    fprintf(f, "  \"sizeof_curandState\": %d,\n  \"cases\": [\n", (int)sizeof(curandState));
    for (int c = 0; c < n_cases; c++) {
        curandStateXORWOW_t st;
        curand_init(cases[c].seed, cases[c].seq, cases[c].off, &st);
        fprintf(f, "    {\"seed\": %llu, \"subsequence\": %llu, \"offset\": %llu, \"d\": %u, \"v\": [%u, %u, %u, %u, %u],\n",
                cases[c].seed, cases[c].seq, cases[c].off, st.d, st.v[0], st.v[1], st.v[2], st.v[3], st.v[4]);
        fprintf(f, "     \"draws_head\": [");
        unsigned int d499 = 0, d999 = 0, d1999 = 0;
        uint64_t xor_all = 0, sum_all = 0;
        for (int k = 0; k < 2000; k++) {
            unsigned int x = curand(&st);
            if (k < n_head) fprintf(f, "%u%s", x, k + 1 < n_head ? ", " : "");
            if (k == 499) d499 = x;
            if (k == 999) d999 = x;
            if (k == 1999) d1999 = x;
            xor_all ^= (uint64_t)x << (k & 31);
            sum_all += x;
        }
        fprintf(f, "], \"draw_499\": %u, \"draw_999\": %u, \"draw_1999\": %u, \"xor2000\": %llu, \"sum2000\": %llu}%s\n",
                d499, d999, d1999, (unsigned long long)xor_all, (unsigned long long)sum_all,
                c + 1 < n_cases ? "," : "");
    }
    fprintf(f, "  ]\n}\n");
    if (f != stdout) fclose(f);
    return 0;
}
