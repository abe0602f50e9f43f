// src/3_sensitivity_analysis.cpp -- Q3 driver on the B200 engine: pathwise vega, finite differences
// with common random numbers, recalibrated finite differences, 20-seed validation (replaces the
// reference's src/3_sensitivity_analysis.cu main()).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "hw1f_driver.hpp"

using namespace hw1f_drv;

static void validation_20_runs(Engine& eng, const std::vector<float>& P, const std::vector<float>& f, float K)
{
    const int n_runs = 20;
    std::printf("\nstatistical validation via confidence intervals...\nRunning %d independent Monte Carlo simulations...\n", n_runs);
    const uint64_t base = base_time() * 2ull;   // src/3:539
    std::vector<uint64_t> seeds(n_runs);
    for (int r = 0; r < n_runs; ++r) seeds[r] = base + (uint64_t)r * 982451653ull;
    std::vector<float> vega(n_runs);
    float ms = 0.f;
    require(hw1f_vega_pathwise_batch(eng.h, seeds.data(), n_runs, kNPaths, 5.0f, 10.0f, K, P.data(), f.data(), -1, vega.data(), &ms),
            eng.h, "hw1f_vega_pathwise_batch");
    std::printf("  Completed %d/%d runs in %.2f ms (one launch)\n", n_runs, n_runs, ms);
    const RunStats s = run_stats(vega);
    std::printf("\nSTATISTICAL VALIDATION RESULTS\n\nSample Statistics (N = %d runs):\n  Mean Vega:              %.6f\n", n_runs, s.mean);
    std::printf("  Standard Deviation:     %.6f\n  Standard Error:         %.6f\n  Coefficient of Var:     %.4f%%\n\n", s.sd, s.se, s.cv_pct);
    std::printf("95%% Confidence Interval:\n  Lower Bound:            %.6f\n  Upper Bound:            %.6f\n  Margin of Error:        ±%.6f\n",
                s.lo, s.hi, s.moe);
    std::printf("  Relative Width:         ±%.4f%%\n\nSample Distribution:\n  Min:  %.6f\n  Q1:   %.6f\n  Med:  %.6f\n  Q3:   %.6f\n  Max:  %.6f\n",
                100.0f * s.moe / s.mean, *std::min_element(vega.begin(), vega.end()), vega[n_runs / 4], vega[n_runs / 2],
                vega[3 * n_runs / 4], *std::max_element(vega.begin(), vega.end()));
    if (FILE* csv = std::fopen("data/vega_bootstrap.csv", "w")) {
        std::fprintf(csv, "run,vega\n");
        for (int r = 0; r < n_runs; ++r) std::fprintf(csv, "%d,%.8f\n", r + 1, vega[r]);
        std::fclose(csv);
        std::printf("\nSaved data/vega_bootstrap.csv\n");
    }
    if (FILE* st = std::fopen("data/vega_statistics.txt", "w")) {
        std::fprintf(st, "VEGA ESTIMATE STATISTICAL VALIDATION\n=====================================\n\nMonte Carlo Parameters:\n");
        std::fprintf(st, "  Paths per run:     %llu\n  Independent runs:  %d\n  Total samples:     %llu\n\n", (unsigned long long)kNPaths, n_runs,
                     (unsigned long long)(kNPaths * n_runs));
        std::fprintf(st, "Point Estimate:\n  Mean Vega:         %.6f\n\nUncertainty Quantification:\n  Standard Error:    %.6f (%.4f%%)\n", s.mean,
                     s.se, 100.0f * s.se / s.mean);
        std::fprintf(st, "  95%% CI:             [%.6f, %.6f]\n\nValidation:\n  Differences < %.6f are statistically insignificant\n", s.lo, s.hi,
                     2 * s.se);
        std::fprintf(st, "  at the 95%% confidence level (within 2 SE).\n");
        std::fclose(st);
        std::printf("Saved data/vega_statistics.txt\n");
    }
}

static void method_agreement(float pw, float fd, float se)
{
    const float diff = std::fabs(pw - fd), z = diff / se;
    std::printf("Comparing Pathwise vs Finite Difference:\n  Pathwise Vega:      %.6f\n  Finite Diff Vega:   %.6f\n", pw, fd);
    std::printf("  Absolute Diff:      %.6f\n  Relative Diff:      %.4f%%\n\nStatistical Test (H0: methods agree):\n", diff, 100.0f * diff / pw);
    std::printf("  Standard Error:     %.6f\n  Z-score:            %.4f\n  Critical value:     1.96 (95%% CI)\n\n", se, z);
    std::printf(z > 1.96f ? "  Result: SIGNIFICANT DIFFERENCE (p < 0.05)\n" : "  Result: NO SIGNIFICANT DIFFERENCE (p > 0.05)\n");
}

int main()
{
    Engine eng;
    std::printf("---Question 3: Sensitivity Analysis---\n\n");
    const int nm = eng.p.n_mat;
    std::vector<float> P(nm), f(nm);
    load_floats("data/P.bin", P.data(), nm);
    load_floats("data/f.bin", f.data(), nm);
    const float S1 = 5.0f, S2 = 10.0f, K = std::exp(-0.1f), eps = 0.001f;

    // one handle walks the reference's draw windows: pathwise [0,n), FD [n,2n), recalibrated [2n,..)
    Rng rng(base_time(), kNPaths);
    hw1f_vega_result v{};
    std::printf("\n---PATHWISE DERIVATIVE METHOD---\n\nMethod: Simultaneous simulation of r(t) and d(sigma)r(t)\n");
    std::printf("  Option: ZBC(S1=%.1f, S2=%.1f, K=e^-0.1)\n  Paths:  %llu Monte Carlo simulations\n", S1, S2, (unsigned long long)kNPaths);
    require(hw1f_vega_pathwise(eng.h, rng.h, S1, S2, K, P.data(), f.data(), -1, &v), eng.h, "hw1f_vega_pathwise");
    std::printf("Performance Results\nVega:   %.6f  (standard error %.6f)\nComputation:      %.2f ms\nThroughput:       %.2f M paths/sec\n",
                v.vega_pathwise, v.vega_pathwise_se, v.ms_pathwise, ((double)kNPaths / v.ms_pathwise) / 1000.0);

    std::printf("\n");
    if (ask_yes("Run block size optimization sweep? (y/n): "))
        std::printf("\nThe B200 engine uses one fixed launch shape (512 threads, 2 subsequences per thread, 2 blocks/SM);\n"
                    "the reference's block-size sweep tunes its own kernel and has no counterpart here.\n");

    std::printf("\nFINITE DIFFERENCE APPROXIMATION\n\n");
    require(hw1f_vega_fd(eng.h, rng.h, S1, S2, K, P.data(), f.data(), eps, v.n_steps_S1, &v), eng.h, "hw1f_vega_fd");
    std::printf("  ZBC( sigma - eps) = %.8f\n  ZBC(sigma + eps) = %.8f\n  FD Vega  = %.6f   (%.2f ms, both bumps in one launch)\n",
                v.price_minus, v.price_plus, v.vega_fd, v.ms_fd);

    std::printf("\n Running finite differences with market data recalibration for theoretical accuracy.\n\n");
    require(hw1f_vega_fd_recalibrated(eng.h, rng.h, S1, S2, K, eps, v.n_steps_S1, &v), eng.h, "hw1f_vega_fd_recalibrated");
    std::printf("Computing at sigma - epsilon = %.4f...\n  Price = %.8f\nComputing at sigma + epsilon = %.4f...\nPrice = %.8f\n\n",
                eng.p.sigma - eps, v.price_minus_recal, eng.p.sigma + eps, v.price_plus_recal);
    std::printf("recalibrated Vega: %.6f   (%.2f ms)\n\n", v.vega_fd_recal, v.ms_fd_recal);

    int c;
    while ((c = std::getchar()) != '\n' && c != EOF) {}
    if (ask_yes("Run statistical validation (20 runs)? (y/n): ")) {
        validation_20_runs(eng, P, f, K);
        method_agreement(v.vega_pathwise, v.vega_fd, 0.000089f);   // hard-coded SE of src/3:747
    }

    const float mc = v.vega_pathwise, fd = v.vega_fd, fdr = v.vega_fd_recal;
    std::printf("Pathwise:        %.6f\nFD (no recalibration):        %.6f  (%.2f%% diff)\nFD (recalibrated):    %.6f  (%.2f%% diff)\n", mc, fd,
                100.0 * std::fabs(mc - fd) / std::fabs(mc), fdr, 100.0 * std::fabs(mc - fdr) / std::fabs(mc));
    std::printf(std::fabs(mc - fdr) > std::fabs(mc - fd) ? "\nRecalibration would make it worse/add no significant benefit.\n"
                                                         : "\nRecalibration would help reduce model inconsistency.\n");
    const float abs_diff = std::fabs(mc - fd), rel = 100.0f * abs_diff / std::fabs(fd);
    std::printf("\nCOMPARATIVE ANALYSIS\n\n--- Vega Estimates ---\n  Pathwise Derivative (MC):   %.6f\n  Finite Difference (FD):     %.6f\n\n", mc, fd);
    std::printf("--- Difference Analysis ---\n  Absolute Difference:        %.6f\n  Relative Difference:        %.2f%%\n\n--- Validation ---\n", abs_diff, rel);
    std::printf("  Sign Check:                 %s\n", (mc > 0 && fd > 0) ? "PASS" : "FAIL");
    std::printf("  Magnitude Check:            %s\n", (mc > 0.05f && mc < 0.5f && fd > 0.05f && fd < 0.5f) ? "PASS" : "FAIL");
    std::printf(rel < 10.0f ? "Agreement Level: < 10%% difference\n" : rel < 25.0f ? "Agreement Level: < 25%% difference\n"
                : rel < 50.0f ? "Agreement Level: < 50%% difference\n" : "Agreement Level: > 50%% difference\n");

    {
        JsonDoc js("data/q3_results.json", "Q3: Sensitivity Analysis", eng.p);
        if (js) std::fprintf(js.file(), "  \"results\": {\n    \"sensitivity_mc\": %.6f,\n    \"sensitivity_fd\": %.6f,\n    \"abs_diff\": %.2e\n  }\n", mc, fd,
                             abs_diff);
    }
    if (FILE* s = summary_section("data/summary.txt", "Q3: SENSITIVITY ANALYSIS")) {
        std::fprintf(s, "  Sens (MC): %.6f\n  Sens (FD): %.6f\n", mc, fd);
        std::fclose(s);
    }
    return 0;
}
