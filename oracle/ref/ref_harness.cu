/*
 * oracle/ref/ref_harness.cu -- TEST INFRASTRUCTURE (reference arm), not product code.
 *
 * Drives the UNMODIFIED reference kernels and host estimators with pinned seeds.  The reference
 * sources are compiled from where they lie (REF_ROOT, normally /root/reference) by #include;
 * nothing is copied.  src/2 and src/3 carry their own main(), which is renamed on inclusion so
 * that their kernels (recover_theta, simulate_sensitivity) and host functions
 * (run_finite_difference, run_finite_difference_recalibrated, ...) can be called from here.
 *
 *   ref_harness parity <seed> <out.json>
 *       Q1 curve (seed), theta recovery, ZBC+CV moments (seed+54321, src/2:128), pathwise vega,
 *       FD and recalibrated FD (seed, src/3:713) with the reference's own draw bookkeeping,
 *       32 sample paths, and raw curandState words of a few paths after init_rng.
 *   ref_harness bench <q1|q2|q3> <steps> <warmup> <out.json>
 *       steady-state CUDA-event timings of the reference path: kernel only (the reference's own
 *       published metric, src/1:64-71) and "workload" = init_rng + kernel + epilogue + D2H.
 *   ref_harness workload <q3seq|zbc20|vega20> <steps> <warmup> <out.json>
 *       wall-clock of the reference's own HOST functions (stdout silenced): q3seq = init_rng +
 *       run_sensitivity_mc + run_finite_difference + run_finite_difference_recalibrated (main() of
 *       src/3); zbc20 / vega20 = the 20-run validations of src/2:210-468 and src/3:527-654
 *       (they write into ./data, so run from a scratch directory).
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define main ref_q2_main
#include "src/2_option_pricing.cu"
#undef main
#define main ref_q3_main
#include "src/3_sensitivity_analysis.cu"
#undef main

static void die(const char* m) { fprintf(stderr, "ref_harness: %s\n", m); exit(2); }
static void ck(const char* what)
{
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "ref_harness: %s: %s\n", what, cudaGetErrorString(e)); exit(3); }
}
static void jarr(FILE* f, const char* name, const float* a, int n, const char* tail)
{
    fprintf(f, "  \"%s\": [", name);
    for (int i = 0; i < n; i++) fprintf(f, "%.9g%s", a[i], i + 1 < n ? ", " : "");
    fprintf(f, "]%s\n", tail);
}

struct Q1Out { float P[N_MAT], f[N_MAT]; float sim_ms; };

/* main() of src/1_bond_pricing.cu:38-83 with a pinned seed */
static void run_q1(unsigned long seed, curandState* d_states, Q1Out* out)
{
    float *d_P_sum, *d_P, *d_f;
    cudaMalloc(&d_P_sum, N_MAT * sizeof(float));
    cudaMalloc(&d_P, N_MAT * sizeof(float));
    cudaMalloc(&d_f, N_MAT * sizeof(float));
    cudaMemset(d_P_sum, 0, N_MAT * sizeof(float));
    compute_constants();
    init_rng<<<NB, NTPB>>>(d_states, seed);
    ck("init_rng");
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    simulate_zcb<<<NB, NTPB>>>(d_P_sum, d_states);
    cudaEventRecord(b);
    ck("simulate_zcb");
    cudaEventElapsedTime(&out->sim_ms, a, b);
    compute_average_and_forward<<<1, 128>>>(d_P, d_f, d_P_sum, N_MAT, 2 * N_PATHS, 1 / H_MAT_SPACING);
    ck("compute_average_and_forward");
    cudaMemcpy(out->P, d_P, N_MAT * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(out->f, d_f, N_MAT * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_P_sum); cudaFree(d_P); cudaFree(d_f);
    cudaEventDestroy(a); cudaEventDestroy(b);
}

static int do_parity(unsigned long seed, const char* path)
{
    FILE* js = fopen(path, "w");
    if (!js) die("cannot open output");
    curandState* d_states;
    cudaMalloc(&d_states, N_PATHS * sizeof(curandState));

    /* ---- raw states after init_rng ---- */
    init_rng<<<NB, NTPB>>>(d_states, seed);
    ck("init_rng");
    const int probe_paths[] = {0, 1, 5, 1023, 1024, 65535, 1048575};
    fprintf(js, "{\n  \"seed\": %lu,\n  \"n_paths\": %d,\n  \"states\": [\n", seed, N_PATHS);
    for (int i = 0; i < 7; i++) {
        curandState h;
        cudaMemcpy(&h, d_states + probe_paths[i], sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(js, "    {\"path\": %d, \"d\": %u, \"v\": [%u, %u, %u, %u, %u]}%s\n", probe_paths[i], h.d, h.v[0], h.v[1],
                h.v[2], h.v[3], h.v[4], i < 6 ? "," : "");
    }
    fprintf(js, "  ],\n");

    /* ---- Q1 ---- */
    Q1Out q1;
    run_q1(seed, d_states, &q1);
    jarr(js, "P", q1.P, N_MAT, ",");
    jarr(js, "f", q1.f, N_MAT, ",");
    fprintf(js, "  \"q1_sim_ms\": %.6f,\n", q1.sim_ms);

    /* simulate_paths_show continues from the advanced states (src/1:163) */
    {
        const int n_show = 32;
        float* d_show;
        std::vector<float> h_show(n_show * (N_STEPS + 1));
        cudaMalloc(&d_show, h_show.size() * sizeof(float));
        simulate_paths_show<<<1, n_show>>>(d_show, d_states, n_show);
        ck("simulate_paths_show");
        cudaMemcpy(h_show.data(), d_show, h_show.size() * sizeof(float), cudaMemcpyDeviceToHost);
        cudaFree(d_show);
        /* keep the fixture small: full first two paths + every 50th value of the rest */
        jarr(js, "r_path0", h_show.data(), N_STEPS + 1, ",");
        jarr(js, "r_path31", h_show.data() + 31 * (N_STEPS + 1), N_STEPS + 1, ",");
    }

    float *d_P_market, *d_f_market;
    load_market_data_to_device(q1.P, q1.f, &d_P_market, &d_f_market);

    /* ---- Q2a: recover_theta (src/2:70-102) ---- */
    {
        float *d_rec, *d_orig, *d_T, h_rec[N_MAT], h_orig[N_MAT], h_T[N_MAT];
        cudaMalloc(&d_rec, N_MAT * sizeof(float));
        cudaMalloc(&d_orig, N_MAT * sizeof(float));
        cudaMalloc(&d_T, N_MAT * sizeof(float));
        recover_theta<<<1, N_MAT>>>(d_f_market, d_rec, d_orig, d_T, N_MAT);
        ck("recover_theta");
        cudaMemcpy(h_rec, d_rec, sizeof(h_rec), cudaMemcpyDeviceToHost);
        cudaMemcpy(h_orig, d_orig, sizeof(h_orig), cudaMemcpyDeviceToHost);
        cudaMemcpy(h_T, d_T, sizeof(h_T), cudaMemcpyDeviceToHost);
        jarr(js, "theta_rec", h_rec, N_MAT, ",");
        jarr(js, "theta_orig", h_orig, N_MAT, ",");
        cudaFree(d_rec); cudaFree(d_orig); cudaFree(d_T);
    }

    /* ---- Q2b: run_ZBC_control_variate (src/2:107-175) with seed + 54321 ---- */
    {
        const float S1 = 5.0f, S2 = 10.0f, K = expf(-0.1f);
        float* d_m;
        cudaMalloc(&d_m, 5 * sizeof(float));
        cudaMemset(d_m, 0, 5 * sizeof(float));
        init_rng<<<NB, NTPB>>>(d_states, seed + 54321);
        ck("init_rng q2");
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        simulate_ZBC_control_variate<<<NB, NTPB>>>(d_m, d_m + 1, d_m + 2, d_m + 3, d_m + 4, d_states, S1, S2, K,
                                                   d_P_market, d_f_market);
        cudaEventRecord(b);
        ck("simulate_ZBC_control_variate");
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        float m[5];
        cudaMemcpy(m, d_m, sizeof(m), cudaMemcpyDeviceToHost);
        const int N_total = 2 * N_PATHS;
        const float mean_ZBC = m[0] / N_total, mean_control = m[1] / N_total;
        const float E_Y2 = m[3] / N_total;
        const float var_control = E_Y2 - mean_control * mean_control;
        const float cov = m[4] / N_total - mean_ZBC * mean_control;
        const float beta = cov / var_control;
        const float adjusted = mean_ZBC - beta * (mean_control - q1.P[N_MAT - 1]);
        const float var_ZBC = m[2] / N_total - mean_ZBC * mean_ZBC;
        const float corr = cov / sqrtf(var_ZBC * var_control);
        jarr(js, "zbc_moments", m, 5, ",");
        fprintf(js, "  \"zbc_mean_X\": %.9g, \"zbc_mean_Y\": %.9g, \"zbc_beta\": %.9g, \"zbc_price_cv\": %.9g, \"zbc_corr\": %.9g, \"zbc_sim_ms\": %.6f,\n",
                mean_ZBC, mean_control, beta, adjusted, corr, ms);
        cudaFree(d_m);
        cudaEventDestroy(a); cudaEventDestroy(b);
    }

    /* ---- Q3: main() of src/3:711-738 with a pinned seed, the reference's own host functions ---- */
    {
        init_rng<<<NB, NTPB>>>(d_states, seed);
        ck("init_rng q3");
        compute_constants();
        float sens_mc = 0, sens_fd = 0, sens_fd_recal = 0;
        run_sensitivity_mc(d_P_market, d_f_market, d_states, &sens_mc);
        run_finite_difference(d_P_market, d_f_market, d_states, &sens_fd);
        run_finite_difference_recalibrated(d_states, &sens_fd_recal);
        ck("q3");
        fprintf(js, "  \"vega_pathwise\": %.9g, \"vega_fd\": %.9g, \"vega_fd_recal\": %.9g,\n", sens_mc, sens_fd,
                sens_fd_recal);
    }
    {
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, 0);
        fprintf(js, "  \"device\": \"%s\"\n}\n", prop.name);
    }
    fclose(js);
    cudaFree(d_states); cudaFree(d_P_market); cudaFree(d_f_market);
    return 0;
}

static int do_bench(const char* which, int steps, int warmup, const char* path)
{
    curandState* d_states;
    cudaMalloc(&d_states, N_PATHS * sizeof(curandState));
    compute_constants();
    Q1Out q1;
    run_q1(1234, d_states, &q1);
    float *d_P_market, *d_f_market;
    load_market_data_to_device(q1.P, q1.f, &d_P_market, &d_f_market);
    float *d_buf, *d_P, *d_f;
    cudaMalloc(&d_buf, N_MAT * sizeof(float));
    cudaMalloc(&d_P, N_MAT * sizeof(float));
    cudaMalloc(&d_f, N_MAT * sizeof(float));
    float* h_pinned;
    cudaMallocHost(&h_pinned, 2 * N_MAT * sizeof(float));
    const float S1 = 5.0f, S2 = 10.0f, K = expf(-0.1f);
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    double kernel_ms = 0, work_ms = 0, init_ms = 0;
    double path_steps = 0;
    for (int it = 0; it < warmup + steps; it++) {
        const unsigned long seed = 1000 + it;
        cudaEventRecord(e0);
        /* what every reference driver does per run: seed, simulate, finalise, copy back */
        if (!strcmp(which, "q1")) compute_constants();   /* H2D of the model tables, src/1:49 */
        init_rng<<<NB, NTPB>>>(d_states, seed);
        cudaEventRecord(e1);
        if (!strcmp(which, "q1")) {
            cudaMemsetAsync(d_buf, 0, N_MAT * sizeof(float));
            simulate_zcb<<<NB, NTPB>>>(d_buf, d_states);
            cudaEventRecord(e2);
            compute_average_and_forward<<<1, 128>>>(d_P, d_f, d_buf, N_MAT, 2 * N_PATHS, 1 / H_MAT_SPACING);
            cudaMemcpyAsync(h_pinned, d_P, N_MAT * sizeof(float), cudaMemcpyDeviceToHost);
            cudaMemcpyAsync(h_pinned + N_MAT, d_f, N_MAT * sizeof(float), cudaMemcpyDeviceToHost);
            path_steps = 2.0 * N_PATHS * N_STEPS;
        } else if (!strcmp(which, "q2")) {
            cudaMemsetAsync(d_buf, 0, 5 * sizeof(float));
            simulate_ZBC_control_variate<<<NB, NTPB>>>(d_buf, d_buf + 1, d_buf + 2, d_buf + 3, d_buf + 4, d_states, S1,
                                                       S2, K, d_P_market, d_f_market);
            cudaEventRecord(e2);
            cudaMemcpyAsync(h_pinned, d_buf, 5 * sizeof(float), cudaMemcpyDeviceToHost);
            path_steps = 2.0 * N_PATHS * 500;
        } else {
            cudaMemsetAsync(d_buf, 0, sizeof(float));
            simulate_sensitivity<<<NB, NTPB>>>(d_buf, d_states, S1, S2, K, d_P_market, d_f_market);
            cudaEventRecord(e2);
            cudaMemcpyAsync(h_pinned, d_buf, sizeof(float), cudaMemcpyDeviceToHost);
            path_steps = 2.0 * N_PATHS * 500;   /* r and d(r)/d(sigma): two processes per path */
        }
        cudaEventRecord(e3);
        ck("bench step");
        if (it >= warmup) {
            float a, b, c;
            cudaEventElapsedTime(&a, e1, e2);
            cudaEventElapsedTime(&b, e0, e3);
            cudaEventElapsedTime(&c, e0, e1);
            kernel_ms += a; work_ms += b; init_ms += c;
        }
    }
    FILE* js = fopen(path, "w");
    if (!js) die("cannot open output");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    fprintf(js, "{\"workload\": \"%s\", \"steps\": %d, \"warmup\": %d, \"path_steps_per_step\": %.0f, "
                "\"kernel_ms_per_step\": %.6f, \"init_rng_ms_per_step\": %.6f, \"workload_ms_per_step\": %.6f, "
                "\"device\": \"%s\"}\n",
            which, steps, warmup, path_steps, kernel_ms / steps, init_ms / steps, work_ms / steps, prop.name);
    fclose(js);
    return 0;
}

static int do_workload(const char* which, int steps, int warmup, const char* path)
{
    curandState* d_states;
    cudaMalloc(&d_states, N_PATHS * sizeof(curandState));
    compute_constants();
    Q1Out q1;
    run_q1(1234, d_states, &q1);
    float *d_P_market, *d_f_market;
    load_market_data_to_device(q1.P, q1.f, &d_P_market, &d_f_market);
    FILE* keep = fopen(path, "w");
    if (!keep) die("cannot open output");
    fflush(stdout);
    if (!freopen("/dev/null", "w", stdout)) die("freopen");
    double total = 0;
    float v0 = 0, v1 = 0, v2 = 0;
    for (int it = 0; it < warmup + steps; it++) {
        cudaDeviceSynchronize();
        const auto t0 = std::chrono::steady_clock::now();
        if (!strcmp(which, "q3seq")) {
            init_rng<<<NB, NTPB>>>(d_states, 1000 + it);
            compute_constants();
            run_sensitivity_mc(d_P_market, d_f_market, d_states, &v0);
            run_finite_difference(d_P_market, d_f_market, d_states, &v1);
            run_finite_difference_recalibrated(d_states, &v2);
        } else if (!strcmp(which, "zbc20")) {
            run_zbc_statistical_validation(d_P_market, d_f_market, q1.P[N_MAT - 1]);
        } else {
            run_statistical_validation(d_P_market, d_f_market);
        }
        cudaDeviceSynchronize();
        const auto t1 = std::chrono::steady_clock::now();
        if (it >= warmup) total += std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
    ck("workload");
    fprintf(keep, "{\"workload\": \"%s\", \"steps\": %d, \"warmup\": %d, \"wall_ms_per_step\": %.6f, "
                  "\"last\": [%.9g, %.9g, %.9g]}\n", which, steps, warmup, total / steps, v0, v1, v2);
    fclose(keep);
    return 0;
}

int main(int argc, char** argv)
{
    if (argc >= 6 && !strcmp(argv[1], "workload")) return do_workload(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5]);
    if (argc >= 4 && !strcmp(argv[1], "parity")) return do_parity(strtoul(argv[2], NULL, 10), argv[3]);
    if (argc >= 6 && !strcmp(argv[1], "bench")) return do_bench(argv[2], atoi(argv[3]), atoi(argv[4]), argv[5]);
    fprintf(stderr, "usage: ref_harness parity <seed> <out.json> | bench <q1|q2|q3> <steps> <warmup> <out.json>\n");
    return 1;
}
