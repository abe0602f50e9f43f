// src/benchmark_reductions.cpp -- reduction-strategy benchmark on the B200 engine (replaces the
// reference's src/benchmark_reductions.cu main()): the reference's three strategies re-expressed on
// the engine's stateless streams plus the engine's deterministic two-level tree as a fourth method.
#include <cmath>
#include <cstdio>
#include <vector>

#include "hw1f_driver.hpp"

using namespace hw1f_drv;

struct Row { const char* method; float ms, mpaths, price; };

int main()
{
    Engine eng;
    std::printf("\n%sREDUCTION METHOD PERFORMANCE BENCHMARK\n%s\n", rule(), rule());
    const int nm = eng.p.n_mat;
    std::vector<float> P(nm), f(nm);
    load_floats("data/P.bin", P.data(), nm);
    load_floats("data/f.bin", f.data(), nm);
    const float S1 = 5.0f, S2 = 10.0f, K = std::exp(-0.1f);
    std::printf("Test Parameters:\n  Option: ZBC(S1=%.1f, S2=%.1f, K=%.6f)\n  Paths: %llu (x2 antithetic = %llu effective)\n", S1, S2, K,
                (unsigned long long)kNPaths, (unsigned long long)(2 * kNPaths));
    std::printf("  Block config: 512 threads/block, 2 subsequences per thread\n  Number of benchmark runs: 5 (average taken)\n\nRunning benchmarks...\n\n");

    const char* names[4] = {"Naive (direct atomicAdd)", "Shared Memory Reduction", "Warp+Block Optimized", "Deterministic two-level tree"};
    Rng rng(base_time(), kNPaths);   // one stream set shared by all methods, advanced by every launch
    std::vector<Row> rows;
    for (int m = 0; m < 4; ++m) {
        float ms = 0.f, price = 0.f;
        std::printf("  Benchmarking %s...\n", names[m]);
        require(hw1f_reduction_bench(eng.h, rng.h, m, S1, S2, K, P.data(), f.data(), -1, 2, 5, &ms, &price), eng.h, "hw1f_reduction_bench");
        const float thr = (float)((2.0 * (double)kNPaths / ms) / 1000.0);
        std::printf("    Time: %.3f ms | Throughput: %.2f M paths/sec | Price: %.8f\n", ms, thr, price);
        rows.push_back({names[m], ms, thr, price});
    }
    std::printf("\n%sBENCHMARK SUMMARY\n%s\n%-30s | %10s | %15s\n", rule(), rule(), "Method", "Time (ms)", "Throughput (M/s)");
    for (const Row& r : rows) std::printf("%-30s | %10.3f | %15.2f  (%.2fx)\n", r.method, r.ms, r.mpaths, rows[0].ms / r.ms);
    std::printf("\n%sVALIDATION\n%s\nPrice consistency (each method runs on the next window of the streams):\n", rule(), rule());
    for (size_t i = 1; i < rows.size(); ++i)
        std::printf("  %-28s vs naive: %.2e (relative: %.4f%%)\n", rows[i].method, std::fabs(rows[0].price - rows[i].price),
                    100.0f * std::fabs(rows[0].price - rows[i].price) / rows[0].price);

    if (FILE* js = std::fopen("data/benchmark_reductions.json", "w")) {
        std::fprintf(js, "{\n  \"benchmark\": \"Reduction Methods Performance\",\n  \"parameters\": {\n    \"N_PATHS\": %llu,\n    \"NTPB\": %d,\n    \"NB\": %llu,\n",
                     (unsigned long long)kNPaths, 512, (unsigned long long)(kNPaths / 1024));
        std::fprintf(js, "    \"S1\": %.1f,\n    \"S2\": %.1f,\n    \"K\": %.6f\n  },\n  \"results\": [\n", S1, S2, K);
        for (size_t i = 0; i < rows.size(); ++i)
            std::fprintf(js, "    {\n      \"method\": \"%s\",\n      \"time_ms\": %.3f,\n      \"throughput_Mpaths_per_sec\": %.2f,\n      \"price\": %.8f\n    }%s\n",
                         rows[i].method, rows[i].ms, rows[i].mpaths, rows[i].price, i + 1 < rows.size() ? "," : "");
        std::fprintf(js, "  ]\n}\n");
        std::fclose(js);
        std::printf("\nSaved data/benchmark_reductions.json\n");
    }
    std::printf("\n%sBENCHMARK COMPLETE\n%s\n", rule(), rule());
    return 0;
}
