// src/1_bond_pricing.cpp -- Q1 driver on the B200 engine (replaces the reference's
// src/1_bond_pricing.cu main(): same console sections, same data/ files, same make target).
#include <cmath>
#include <cstdio>
#include <vector>

#include "hw1f_driver.hpp"

using namespace hw1f_drv;

int main()
{
    std::printf("ZERO COUPON BOND PRICING\n");
    Engine eng;
    const hw1f_params& p = eng.p;
    const int nm = p.n_mat, stride = eng.c.save_stride;
    std::printf("Parameters:\n  N_PATHS = %llu (x2 antithetic = %llu effective)\n", (unsigned long long)kNPaths,
                (unsigned long long)(2 * kNPaths));
    std::printf("  N_STEPS = %d, N_MAT = %d, T = %.1f years\n  a = %.2f, sigma = %.2f, r0 = %.4f\n\n", p.n_steps, nm,
                p.T_final, p.a, p.sigma, p.r0);

    Rng rng(base_time(), kNPaths);   // init_rng(states, time(NULL))
    std::vector<float> P(nm), f(nm), se(nm);
    float sim_ms = 0.f;
    std::printf("Running Monte Carlo simulation...\n");
    int gpus = 1;
    if (const char* s = std::getenv("HW_GPUS")) gpus = std::atoi(s);
    if (gpus > 1) {
        // same path set, sharded by subsequence range over `gpus` devices + one NCCL all-reduce
        hw1f_multi* m = nullptr;
        if (hw1f_multi_create(gpus, &m) != HW1F_OK) { std::fprintf(stderr, "hw1f_multi_create failed\n"); return 1; }
        uint64_t seed = 0;
        hw1f_rng_info(rng.h, &seed, nullptr, nullptr);
        int mode = HW1F_MODE_DECOMPOSED;
        hw1f_engine_get_mode(eng.h, &mode);
        hw1f_multi_set_mode(m, mode);
        if (hw1f_multi_set_model(m, &p) != HW1F_OK ||
            hw1f_multi_bond_curve(m, seed, kNPaths, 0, P.data(), f.data(), se.data(), &sim_ms) != HW1F_OK) {
            std::fprintf(stderr, "multi-GPU run failed: %s\n", hw1f_multi_last_error(m));
            return 1;
        }
        int used = 0;
        hw1f_multi_device_count(m, &used);
        std::printf("(sharded over %d GPUs)\n", used);
        hw1f_multi_destroy(m);
        hw1f_rng_seek(rng.h, (uint64_t)p.n_steps);   // where the single-GPU run would have left the streams
    } else {
        require(hw1f_bond_curve(eng.h, rng.h, P.data(), f.data(), se.data(), &sim_ms), eng.h, "hw1f_bond_curve");
    }
    std::printf("Simulation complete\n\nRESULTS\nT (years)    P(0,T)         f(0,T)\n");
    for (int i = 0; i < nm; i += stride)
        std::printf("%5.1f        %.6f       %7.4f%%\n", i * eng.c.mat_spacing, P[i], f[i] * 100.0f);

    std::printf("\nChecks\n");
    std::printf("P(0,0) = 1.0:      %.6f %s\n", P[0], (P[0] > 0.99f && P[0] < 1.01f) ? "OK" : "ERROR");
    std::printf("P(0,10) ~ 0.87:    %.6f %s\n", P[nm - 1], (P[nm - 1] > 0.3f && P[nm - 1] < 0.9f) ? "OK" : "ERROR");
    std::printf("f(0,0) ~ 1.2%%:     %.4f%% %s\n", f[0] * 100.0f, (f[0] > 0.01f && f[0] < 0.02f) ? "OK" : "ERROR");
    std::printf("P(0,10) standard error: %.2e\n", se[nm - 1]);

    const double n_eff = 2.0 * (double)kNPaths;
    std::printf("\nPerformance\nSimulation time: %.2f ms\nEffective paths: %.0f\nThroughput: %.2f M paths/sec\n", sim_ms,
                n_eff, (n_eff / sim_ms) / 1000.0);

    summary_start("data/summary.txt", p);
    std::printf("\nSaving Results\n");
    save_floats("data/P.bin", P.data(), nm);
    save_floats("data/f.bin", f.data(), nm);
    {
        JsonDoc js("data/q1_results.json", "Q1: Zero-Coupon Bond Pricing", p);
        if (js) {
            js.array("P", P.data(), nm, true);
            js.array("f", f.data(), nm, true);
            js.performance(sim_ms, n_eff, true);
            std::fprintf(js.file(), "  \"validation\": {\n    \"P_0_0\": %.8f,\n    \"P_0_10\": %.8f,\n    \"f_0_0\": %.8f\n  }\n",
                         P[0], P[nm - 1], f[0]);
        }
    }
    csv_series("data/P_curve.csv", "P(0 T)", P.data(), nm, eng.c.mat_spacing);
    csv_series("data/f_curve.csv", "f(0 T)", f.data(), nm, eng.c.mat_spacing);
    if (FILE* s = summary_section("data/summary.txt", "Q1: ZERO-COUPON BOND PRICING")) {
        std::fprintf(s, "\nKey Results:\n  P(0,0) = %.8f (expected: 1.0)\n  P(0,10) = %.8f\n  f(0,0) = %.4f%% (expected: ~1.2%%)\n",
                     P[0], P[nm - 1], f[0] * 100.0f);
        std::fprintf(s, "\nPerformance:\n  Simulation time: %.2f ms\n  Throughput: %.2f M paths/sec\n", sim_ms,
                     (n_eff / sim_ms) / 1000.0);
        std::fclose(s);
    }

    // 32 sample trajectories; like the reference they continue the streams after the curve run
    const int n_show = 32;
    std::vector<float> paths((size_t)n_show * (p.n_steps + 1));
    require(hw1f_sample_paths(eng.h, rng.h, n_show, paths.data()), eng.h, "hw1f_sample_paths");
    save_floats("data/r_paths.bin", paths.data(), (int)paths.size());
    return 0;
}
