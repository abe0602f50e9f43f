// tools/api_overhead.cpp -- host overhead of the C ABI measured from C++ (no Python marshalling in the way):
// wall clock per public call next to the CUDA-event time the call reports for its GPU part.
//   make api_overhead && ./bin/api_overhead gpurun_out/api_overhead_c.json
#include <chrono>
#include <cstdio>
#include <vector>

#include "hw1f_driver.hpp"

using namespace hw1f_drv;
using clk = std::chrono::steady_clock;

static std::FILE* g_out = stdout;

template <class F>
static void timed(const char* name, int steps, F&& fn, bool last = false)
{
    for (int i = 0; i < 5; ++i) fn(i);
    double ev = 0.0;
    const auto t0 = clk::now();
    for (int i = 0; i < steps; ++i) ev += fn(100 + i);
    const double wall = std::chrono::duration<double, std::milli>(clk::now() - t0).count() / steps;
    std::fprintf(g_out, " \"%s\": {\"wall_ms\": %.4f, \"event_ms\": %.4f, \"host_overhead_ms\": %.4f}%s\n", name, wall, ev / steps,
                wall - ev / steps, last ? "" : ",");
}

int main(int argc, char** argv)
{
    if (argc > 1 && !(g_out = std::fopen(argv[1], "w"))) { std::perror(argv[1]); return 1; }
    std::FILE* keep = g_out;
    Engine eng;   // prints "Using GPU n" to stdout
    const int nm = eng.p.n_mat;
    std::vector<float> P(nm), f(nm), se(nm);
    {
        Rng r(1234, kNPaths);
        require(hw1f_bond_curve(eng.h, r.h, P.data(), f.data(), se.data(), nullptr), eng.h, "hw1f_bond_curve");
    }
    const float S1 = 5.0f, S2 = 10.0f, K = std::exp(-0.1f);
    std::vector<float> P2(nm), f2(nm);
    std::fprintf(keep, "{\n");
    timed("hw1f_bond_curve", 200, [&](int i) {
        Rng r(i, kNPaths); float ms = 0.f;
        require(hw1f_bond_curve(eng.h, r.h, P2.data(), f2.data(), se.data(), &ms), eng.h, "curve"); return (double)ms; });
    timed("hw1f_set_model + hw1f_bond_curve (bench e2e step)", 200, [&](int i) {
        Rng r(i, kNPaths); float ms = 0.f;
        require(hw1f_set_model(eng.h, &eng.p), eng.h, "set_model");
        require(hw1f_bond_curve(eng.h, r.h, P2.data(), f2.data(), se.data(), &ms), eng.h, "curve"); return (double)ms; });
    timed("hw1f_zbc_cv", 200, [&](int i) {
        Rng r(i, kNPaths); float ms = 0.f; hw1f_zbc_result z;
        require(hw1f_zbc_cv(eng.h, r.h, S1, S2, K, P.data(), f.data(), 500, &z, &ms), eng.h, "zbc"); return (double)ms; });
    timed("hw1f_vega_pathwise", 200, [&](int i) {
        Rng r(i, kNPaths); hw1f_vega_result v;
        require(hw1f_vega_pathwise(eng.h, r.h, S1, S2, K, P.data(), f.data(), 500, &v), eng.h, "pw"); return (double)v.ms_pathwise; });
    timed("hw1f_vega_fd", 200, [&](int i) {
        Rng r(i, kNPaths); hw1f_vega_result v;
        require(hw1f_vega_fd(eng.h, r.h, S1, S2, K, P.data(), f.data(), 0.001f, 500, &v), eng.h, "fd"); return (double)v.ms_fd; });
    timed("hw1f_vega (Q3 sequence)", 100, [&](int i) {
        Rng r(i, kNPaths); hw1f_vega_result v;
        require(hw1f_vega(eng.h, r.h, S1, S2, K, P.data(), f.data(), 0.001f, 500, &v), eng.h, "vega");
        return (double)v.ms_pathwise + v.ms_fd + v.ms_fd_recal; });
    timed("hw1f_fused", 100, [&](int i) {
        Rng r(i, kNPaths); hw1f_zbc_result z; hw1f_vega_result v; float ms = 0.f;
        require(hw1f_fused(eng.h, r.h, S1, S2, K, P.data(), f.data(), 0.001f, 500, P2.data(), f2.data(), se.data(), &z, &v, &ms),
                eng.h, "fused"); return (double)ms; }, true);
    std::fprintf(keep, "}\n");
    if (g_out != stdout) std::fclose(g_out);
    return 0;
}
