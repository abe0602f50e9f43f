// include/hw1f_driver.hpp -- host-side helpers shared by the four drivers in src/.
//
// The drivers are plain C++ (no CUDA): everything numerical happens behind the C ABI of
// include/hw1f.h.  This header provides (1) an RAII view of the engine / RNG handles and (2) the
// output schema of the reference's include/output.cuh (same file names, keys, column headers and
// number formats), which analyze.py consumes.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <string>
#include <vector>

#include "hw1f.h"

namespace hw1f_drv {

inline void require(int status, hw1f_engine* eng, const char* what)
{
    if (status == HW1F_OK) return;
    std::fprintf(stderr, "hw1f error in %s: %s (%s)\n", what, hw1f_status_string(status),
                 eng ? hw1f_last_error(eng) : "-");
    std::exit(1);
}

// seeds follow the reference (time(NULL)-derived) unless HW_SEED pins them for A/B runs
inline uint64_t base_time()
{
    if (const char* s = std::getenv("HW_SEED")) return std::strtoull(s, nullptr, 10);
    return (uint64_t)std::time(nullptr);
}

struct Engine {
    hw1f_engine* h = nullptr;
    hw1f_params p{};
    hw1f_constants c{};
    Engine()
    {
        int dev = -1;   // select_gpu(): most free memory; HW_DEVICE overrides
        if (const char* s = std::getenv("HW_DEVICE")) dev = std::atoi(s);
        require(hw1f_engine_create(dev, &h), nullptr, "hw1f_engine_create");
        if (const char* s = std::getenv("HW_MODE"))   // "reference": the reference's own per-path float order
            hw1f_engine_set_mode(h, (s[0] == 'r' || s[0] == '0') ? HW1F_MODE_REFERENCE_ORDER : HW1F_MODE_DECOMPOSED);
        hw1f_default_params(&p);
        require(hw1f_set_model(h, &p), h, "hw1f_set_model");
        hw1f_get_constants(h, &c);
        int d = 0;
        hw1f_engine_device(h, &d);
        std::printf("Using GPU %d\n\n", d);
    }
    ~Engine() { hw1f_engine_destroy(h); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;
};

struct Rng {
    hw1f_rng* h = nullptr;
    Rng(uint64_t seed, uint64_t n_paths) { require(hw1f_rng_create(seed, 0, n_paths, &h), nullptr, "hw1f_rng_create"); }
    ~Rng() { hw1f_rng_destroy(h); }
    Rng(const Rng&) = delete;
    Rng& operator=(const Rng&) = delete;
};

constexpr uint64_t kNPaths = 1024ull * 1024ull;   // N_PATHS of the reference (common.cuh:16)

// ---- data/P.bin, data/f.bin (raw float32[n]) -----------------------------------------------------
inline void save_floats(const char* path, const float* data, int n)
{
    FILE* f = std::fopen(path, "wb");
    if (!f) { std::printf("Error: Cannot open %s for writing\n", path); std::exit(1); }
    std::fwrite(data, sizeof(float), (size_t)n, f);
    std::fclose(f);
    std::printf("Saved %s (%d floats)\n", path, n);
}

inline void load_floats(const char* path, float* data, int n)
{
    FILE* f = std::fopen(path, "rb");
    if (!f) { std::printf("Error: Cannot open %s\nDid you run Q1 first?\n", path); std::exit(1); }
    const size_t got = std::fread(data, sizeof(float), (size_t)n, f);
    std::fclose(f);
    if ((int)got != n) { std::printf("Error: Expected %d floats, got %zu\n", n, got); std::exit(1); }
    std::printf("Loaded %s (%d floats)\n", path, n);
}

// ---- JSON documents with the reference's header block ----------------------------------------------
class JsonDoc {
public:
    JsonDoc(const char* path, const char* task, const hw1f_params& p) : f_(std::fopen(path, "w")), path_(path)
    {
        if (!f_) { std::printf("Error: Cannot create %s\n", path); return; }
        std::time_t now = std::time(nullptr);
        std::string stamp = std::ctime(&now);
        if (!stamp.empty() && stamp.back() == '\n') stamp.pop_back();
        std::fprintf(f_, "{\n  \"task\": \"%s\",\n  \"timestamp\": \"%s\",\n  \"parameters\": {\n", task, stamp.c_str());
        std::fprintf(f_, "    \"N_PATHS\": %llu,\n    \"N_STEPS\": %d,\n    \"N_MAT\": %d,\n    \"T_FINAL\": %.1f,\n",
                     (unsigned long long)kNPaths, p.n_steps, p.n_mat, p.T_final);
        std::fprintf(f_, "    \"a\": %.2f,\n    \"sigma\": %.2f,\n    \"r0\": %.4f\n  },\n", p.a, p.sigma, p.r0);
    }
    ~JsonDoc()
    {
        if (!f_) return;
        std::fprintf(f_, "}\n");
        std::fclose(f_);
        std::printf("Saved %s\n", path_.c_str());
    }
    explicit operator bool() const { return f_ != nullptr; }
    FILE* file() { return f_; }
    void array(const char* name, const float* data, int n, bool comma)
    {
        std::fprintf(f_, "  \"%s\": [", name);
        for (int i = 0; i < n; ++i) {
            if (i % 10 == 0) std::fprintf(f_, "\n    ");
            std::fprintf(f_, "%.8f%s", data[i], i + 1 < n ? ", " : "");
        }
        std::fprintf(f_, "\n  ]%s\n", comma ? "," : "");
    }
    void performance(float ms, double n_paths, bool comma)
    {
        std::fprintf(f_, "  \"performance\": {\n    \"simulation_time_ms\": %.2f,\n    \"throughput_Mpaths_per_sec\": %.2f\n  }%s\n",
                     ms, (n_paths / ms) / 1000.0, comma ? "," : "");
    }

private:
    FILE* f_;
    std::string path_;
};

// "T,<header>" time series and three-column comparison CSVs
inline void csv_series(const char* path, const char* header, const float* data, int n, float spacing)
{
    FILE* f = std::fopen(path, "w");
    if (!f) { std::printf("Error: Cannot create %s\n", path); return; }
    std::fprintf(f, "T,%s\n", header);
    for (int i = 0; i < n; ++i) std::fprintf(f, "%.4f,%.8f\n", i * spacing, data[i]);
    std::fclose(f);
    std::printf("Saved %s\n", path);
}

inline void csv_three(const char* path, const char* h0, const char* h1, const char* h2, const float* x, const float* y1,
                      const float* y2, int n)
{
    FILE* f = std::fopen(path, "w");
    if (!f) { std::printf("Error: Cannot create %s\n", path); return; }
    std::fprintf(f, "%s,%s,%s\n", h0, h1, h2);
    for (int i = 0; i < n; ++i) std::fprintf(f, "%.4f,%.8f,%.8f\n", x[i], y1[i], y2[i]);
    std::fclose(f);
    std::printf("Saved %s\n", path);
}

// data/summary.txt
inline const char* rule() { return "================================================================================\n"; }

inline void summary_start(const char* path, const hw1f_params& p)
{
    FILE* f = std::fopen(path, "w");
    if (!f) { std::printf("Error: Cannot create %s\n", path); return; }
    std::time_t now = std::time(nullptr);
    std::fprintf(f, "%sHULL-WHITE MODEL SIMULATION RESULTS\n%sGenerated: %s\nParameters:\n", rule(), rule(), std::ctime(&now));
    std::fprintf(f, "  N_PATHS = %llu (x2 antithetic = %llu effective)\n  N_STEPS = %d\n  N_MAT = %d\n",
                 (unsigned long long)kNPaths, (unsigned long long)(2 * kNPaths), p.n_steps, p.n_mat);
    std::fprintf(f, "  T_FINAL = %.1f years\n  a = %.2f, sigma = %.2f, r0 = %.4f\n", p.T_final, p.a, p.sigma, p.r0);
    std::fclose(f);
    std::printf("Initialized %s\n", path);
}

inline FILE* summary_section(const char* path, const char* title)
{
    FILE* f = std::fopen(path, "a");
    if (!f) { std::printf("Error: Cannot open %s\n", path); return nullptr; }
    std::fprintf(f, "\n%s%s\n%s", rule(), title, rule());
    return f;
}

// 20-run statistics in float32, in the reference's operation order (src/2:305-324, src/3:570-589)
struct RunStats { float mean, variance, sd, se, moe, lo, hi, cv_pct; };
inline RunStats run_stats(const std::vector<float>& x)
{
    RunStats s{};
    const int n = (int)x.size();
    for (float v : x) s.mean += v;
    s.mean /= n;
    for (float v : x) { const float d = v - s.mean; s.variance += d * d; }
    s.variance /= (n - 1);
    s.sd = std::sqrt(s.variance);
    s.se = s.sd / std::sqrt((float)n);
    s.moe = 2.093f * s.se;   // t(0.975, 19)
    s.lo = s.mean - s.moe;
    s.hi = s.mean + s.moe;
    s.cv_pct = 100.0f * s.sd / s.mean;
    return s;
}

inline bool ask_yes(const char* prompt)
{
    std::printf("%s", prompt);
    std::fflush(stdout);
    char c = 'n';
    if (std::scanf(" %c", &c) != 1) return false;
    return c == 'y' || c == 'Y';
}

}  // namespace hw1f_drv
