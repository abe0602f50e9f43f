// hw1f_kernels_extra.cuh -- (1) the reduction-strategy benchmark kernels behind
// hw1f_reduction_bench (the comparison points of include/perf_benchmark.cuh:19-197 re-expressed on
// this engine's stateless streams, plus the engine's deterministic tree), and (2) the fused
// curve + ZBC/control + pathwise-tangent pass of the BASELINE.json scaling run.
#pragma once
#include "hw1f_kernels.cuh"

namespace hw1f {

// ---- shared: antithetic simulation to S1 for one scenario, both lanes ------------------------------
struct PairState { float2 r1, r2, I1, I2; };

__device__ __forceinline__ PairState simulate_to_S1(ThreadStreams& t, const float4* __restrict__ drift4,
                                                    const float2* __restrict__ drift2_gl, int n_steps_S1, int lead,
                                                    float r0, float sig_st, float2 e2, float2 hdt2)
{
    PairState st;
    st.r1 = st.r2 = splat(r0);
    st.I1 = st.I2 = splat(0.0f);
    const float2 sgP = splat(sig_st), sgM = splat(-sig_st);
    auto step1 = [&](float2 d, float2 G) {
        hw_step2(st.r1, st.I1, fma2(G, sgP, d), e2, hdt2);
        hw_step2(st.r2, st.I2, fma2(G, sgM, d), e2, hdt2);
    };
    auto pairfn = [&](int pk, float2 ns, float2 nc) {
        const float4 d = drift4[pk];
        step1(make_float2(d.x, d.y), ns);
        step1(make_float2(d.z, d.w), nc);
    };
    const int n_main = max(n_steps_S1 - lead, 0);
    if (lead && n_steps_S1 > 0) {
        float2 ns, nc;
        one_pair(t, ns, nc);
        step1(drift2_gl[0], nc);
    }
    int pair = 0;
    advance_pairs(t, pair, n_main >> 1, pairfn);
    if (n_main & 1) {
        float2 ns, nc;
        one_pair(t, ns, nc);
        const float4 d = drift4[pair];
        step1(make_float2(d.x, d.y), ns);
    }
    return st;
}

// ZBC payoff of both twins of both lanes: x = discount * max(P - K, 0)  (common.cuh:337-353)
__device__ __forceinline__ void zbc_payoffs(const PairState& st, const BondPlan& pl, float K, float2& x1, float2& x2,
                                            float2& c1, float2& c2)
{
    const float2 z1 = mul2(mul2(st.r1, splat(pl.negB)), splat(kLog2e));
    const float2 z2 = mul2(mul2(st.r2, splat(pl.negB)), splat(kLog2e));
    const float2 P1 = mul2(splat(pl.A), make_float2(mufu_ex2(z1.x), mufu_ex2(z1.y)));
    const float2 P2 = mul2(splat(pl.A), make_float2(mufu_ex2(z2.x), mufu_ex2(z2.y)));
    const float2 q1 = mul2(st.I1, splat(-kLog2e)), q2 = mul2(st.I2, splat(-kLog2e));
    const float2 d1 = make_float2(mufu_ex2(q1.x), mufu_ex2(q1.y));
    const float2 d2 = make_float2(mufu_ex2(q2.x), mufu_ex2(q2.y));
    c1 = mul2(P1, d1);
    c2 = mul2(P2, d2);
    const float2 g1 = add2(P1, splat(-K)), g2 = add2(P2, splat(-K));
    x1 = mul2(d1, make_float2(fmaxf(0.0f, g1.x), fmaxf(0.0f, g1.y)));
    x2 = mul2(d2, make_float2(fmaxf(0.0f, g2.x), fmaxf(0.0f, g2.y)));
}

// =================================================================================================
// reduction benchmark: sum of ZBC payoffs, four strategies
// =================================================================================================
// METHOD 0: one float atomicAdd per reference thread (simulate_ZBC_naive, perf_benchmark.cuh:57)
// METHOD 1: shared-memory tree + one atomic per block (simulate_ZBC_shared_memory, :118-127)
// METHOD 2: warp shuffle + block shuffle + one atomic per block (simulate_ZBC_warp_optimized, :183-196)
// METHOD 3: the engine's deterministic tree: no atomics, double partial per block
template <int METHOD>
__global__ void __launch_bounds__(kThreads, 2)
zbc_sum_kernel(StreamGeom g, SeedArgs seeds, ModelDev md, ScenDev sc, const BondPlan* __restrict__ plans,
               int n_steps_S1, int lead, float K, float* __restrict__ sum_f, double* __restrict__ partials)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* win = smem;
    const int n_main = max(n_steps_S1 - lead, 0);
    const int n_slots = (n_main + 1) >> 1;
    float4* drift4 = reinterpret_cast<float4*>(smem + kWinWords);
    __shared__ float s_red[kThreads];
    __shared__ double s_dbl[kWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n_slots; i += kThreads) {
        const int i0 = lead + 2 * i, i1 = i0 + 1;
        const float2 a = sc.drift2[i0];
        const float2 b = (i1 < n_steps_S1) ? sc.drift2[i1] : make_float2(0.0f, 0.0f);
        drift4[i] = make_float4(a.x, a.y, b.x, b.y);
    }
    const BondPlan pl = plans[0];
    const float2 e2 = splat(md.exp_adt), hdt2 = splat(mul_(0.5f, md.dt));
    double block_acc = 0.0;
    for (unsigned long long chunk = blockIdx.x; chunk < g.n_chunks; chunk += gridDim.x) {
        ThreadStreams t = derive_streams(g, seeds, 0, chunk, win);
        const PairState st = simulate_to_S1(t, drift4, sc.drift2, n_steps_S1, lead, md.r0, sc.sig_st, e2, hdt2);
        float2 x1, x2, c1, c2;
        zbc_payoffs(st, pl, K, x1, x2, c1, c2);
        const float2 x = add2(x1, x2);                         // payoff1 + payoff2 per reference thread
        const float xa = t.validA ? x.x : 0.0f, xb = t.validB ? x.y : 0.0f;
        if (METHOD == 0) {
            if (t.validA) atomicAdd(sum_f, xa);
            if (t.validB) atomicAdd(sum_f, xb);
        } else if (METHOD == 1) {
            __syncthreads();
            s_red[tid] = add_(xa, xb);
            __syncthreads();
            for (int s = kThreads / 2; s > 0; s >>= 1) {
                if (tid < s) s_red[tid] += s_red[tid + s];
                __syncthreads();
            }
            if (tid == 0) atomicAdd(sum_f, s_red[0]);
        } else if (METHOD == 2) {
            float v = add_(xa, xb);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            __syncthreads();
            if (lane == 0) s_red[warp] = v;
            __syncthreads();
            if (warp == 0) {
                float w = (lane < kWarps) ? s_red[lane] : 0.0f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
                if (lane == 0) atomicAdd(sum_f, w);
            }
        } else {
            const double w = warp_sum((double)xa + (double)xb);
            __syncthreads();
            if (lane == 0) s_dbl[warp] = w;
            __syncthreads();
            if (tid == 0) {
                double acc = s_dbl[0];
#pragma unroll
                for (int k = 1; k < kWarps; ++k) acc += s_dbl[k];
                block_acc += acc;
            }
        }
    }
    if (METHOD == 3 && tid == 0) partials[blockIdx.x] = block_acc;
}

// =================================================================================================
// fused pass: curve sums on the maturity grid + ZBC/control moments + pathwise tangent at S1,
// all from ONE set of normals (BASELINE.json scaling-run workload)
// =================================================================================================
constexpr int kFusedExtra = 8;     // 5 ZBC moments, sum(v1+v2), sum (v1+v2)^2, sum v1
constexpr int kFusedFdExtra = 10;  // + 5 ZBC moments at sigma-eps and 5 at sigma+eps (FD variant)

// partials[block][2*n_mat + kFusedExtra (+ kFusedFdExtra)] doubles.  Requires even stride, even
// n_steps_S1 that is a multiple of the stride, even normal offset (checked on the host).
// FD = true adds the two bumped-sigma antithetic pairs of run_finite_difference (src/3:400-446) on the
// SAME normals: scm/scp carry sig_st and the shifted drift tables, plans[1], plans[2] their A(S1,S2).
template <bool FD>
__global__ void __launch_bounds__(kThreads, 1)
fused_kernel(StreamGeom g, SeedArgs seeds, ModelDev md, ScenDev sc, ScenDev scm, ScenDev scp,
             const BondPlan* __restrict__ plans, int n_steps_S1, float K, double* __restrict__ partials)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int n_steps = md.n_steps, n_mat = md.n_mat;
    constexpr int kExt = kFusedExtra + (FD ? kFusedFdExtra : 0);
    const int nq = 2 * n_mat + kExt;
    const int n_pairs_tot = n_steps >> 1;
    const int n_pairs_S1 = n_steps_S1 >> 1;
    uint32_t* win = smem;
    float4* drift4 = reinterpret_cast<float4*>(smem + kWinWords);                 // [n_steps/2]
    float4* sdrift4 = drift4 + n_pairs_tot;                                       // [n_steps_S1/2]
    float4* dm4 = sdrift4 + n_pairs_S1;                                           // [n_steps_S1/2] (FD) sigma-eps drift
    float4* dp4 = dm4 + (FD ? n_pairs_S1 : 0);                                    // [n_steps_S1/2] (FD) sigma+eps drift
    double* bacc = reinterpret_cast<double*>(dp4 + (FD ? n_pairs_S1 : 0));        // [nq]
    float* wflt = reinterpret_cast<float*>(bacc + nq);                            // [kWarps][2*n_mat]
    float* cen = wflt + kWarps * 2 * n_mat;                                       // [n_mat]
    __shared__ double wext[kWarps][kExt];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < n_pairs_tot; i += kThreads) drift4[i] = reinterpret_cast<const float4*>(sc.drift2)[i];
    for (int i = tid; i < n_pairs_S1; i += kThreads) {
        sdrift4[i] = reinterpret_cast<const float4*>(sc.sdrift2)[i];
        if (FD) {
            dm4[i] = reinterpret_cast<const float4*>(scm.drift2)[i];
            dp4[i] = reinterpret_cast<const float4*>(scp.drift2)[i];
        }
    }
    for (int k = tid; k < nq; k += kThreads) bacc[k] = 0.0;
    for (int k = tid; k < kWarps * 2 * n_mat; k += kThreads) wflt[k] = 0.0f;
    for (int k = tid; k < n_mat; k += kThreads) cen[k] = sc.center[k];

    const BondPlan pl = plans[0];
    const float2 sgP = splat(sc.sig_st), sgM = splat(-sc.sig_st), ctP = splat(pl.c_t), ctM = splat(-pl.c_t);
    const float2 e2 = splat(md.exp_adt), hdt2 = splat(mul_(0.5f, md.dt));
    const float2 smP = splat(scm.sig_st), smM = splat(-scm.sig_st), spP = splat(scp.sig_st), spM = splat(-scp.sig_st);
    const int half = md.stride >> 1;
    const int m_S1 = n_steps_S1 / md.stride;          // maturity index reached at S1
    const bool writer = (lane & 15) == 0;
    float* const wrow = wflt + warp * 2 * n_mat + ((lane & 16) ? n_mat : 0);

    for (unsigned long long chunk = blockIdx.x; chunk < g.n_chunks; chunk += gridDim.x) {
        ThreadStreams t = derive_streams(g, seeds, 0, chunk, win);
        float2 r1 = splat(md.r0), r2 = splat(md.r0), I1 = splat(0.0f), I2 = splat(0.0f);
        float2 t1 = splat(0.0f), t2 = splat(0.0f), J1 = splat(0.0f), J2 = splat(0.0f);   // tangents and their integrals
        PairState bm, bp;                                                                  // bumped-sigma pairs (FD)
        bm.r1 = bm.r2 = bp.r1 = bp.r2 = splat(md.r0);
        bm.I1 = bm.I2 = bp.I1 = bp.I2 = splat(0.0f);
        const float2 mask = make_float2(t.validA ? 1.0f : 0.0f, t.validB ? 1.0f : 0.0f);
        const bool full = __syncthreads_and(t.validA && t.validB);

        auto pair_tan = [&](int pk, float2 ns, float2 nc) {       // steps up to S1: r and d(r)/d(sigma)
            const float4 d = drift4[pk], sd = sdrift4[pk];
            const float2 da = make_float2(d.x, d.y), db = make_float2(d.z, d.w);
            const float2 sa = make_float2(sd.x, sd.y), sb = make_float2(sd.z, sd.w);
            hw_step2(r1, I1, fma2(ns, sgP, da), e2, hdt2);
            hw_step2(r2, I2, fma2(ns, sgM, da), e2, hdt2);
            hw_step2(t1, J1, fma2(ctP, ns, sa), e2, hdt2);
            hw_step2(t2, J2, fma2(ctM, ns, sa), e2, hdt2);
            hw_step2(r1, I1, fma2(nc, sgP, db), e2, hdt2);
            hw_step2(r2, I2, fma2(nc, sgM, db), e2, hdt2);
            hw_step2(t1, J1, fma2(ctP, nc, sb), e2, hdt2);
            hw_step2(t2, J2, fma2(ctM, nc, sb), e2, hdt2);
            if (FD) {
                const float4 m4 = dm4[pk], p4 = dp4[pk];
                const float2 ma = make_float2(m4.x, m4.y), mb = make_float2(m4.z, m4.w);
                const float2 pa = make_float2(p4.x, p4.y), pb = make_float2(p4.z, p4.w);
                hw_step2(bm.r1, bm.I1, fma2(ns, smP, ma), e2, hdt2);
                hw_step2(bm.r2, bm.I2, fma2(ns, smM, ma), e2, hdt2);
                hw_step2(bp.r1, bp.I1, fma2(ns, spP, pa), e2, hdt2);
                hw_step2(bp.r2, bp.I2, fma2(ns, spM, pa), e2, hdt2);
                hw_step2(bm.r1, bm.I1, fma2(nc, smP, mb), e2, hdt2);
                hw_step2(bm.r2, bm.I2, fma2(nc, smM, mb), e2, hdt2);
                hw_step2(bp.r1, bp.I1, fma2(nc, spP, pb), e2, hdt2);
                hw_step2(bp.r2, bp.I2, fma2(nc, spM, pb), e2, hdt2);
            }
        };
        auto pair_plain = [&](int pk, float2 ns, float2 nc) {     // after S1: curve only
            const float4 d = drift4[pk];
            const float2 da = make_float2(d.x, d.y), db = make_float2(d.z, d.w);
            hw_step2(r1, I1, fma2(ns, sgP, da), e2, hdt2);
            hw_step2(r2, I2, fma2(ns, sgM, da), e2, hdt2);
            hw_step2(r1, I1, fma2(nc, sgP, db), e2, hdt2);
            hw_step2(r2, I2, fma2(nc, sgM, db), e2, hdt2);
        };
        auto save_curve = [&](int m) {
            const float2 a = mul2(I1, splat(-kLog2e)), b = mul2(I2, splat(-kLog2e));
            const float2 p0 = add2(make_float2(mufu_ex2(a.x), mufu_ex2(a.y)), make_float2(mufu_ex2(b.x), mufu_ex2(b.y)));
            float2 dv = add2(p0, splat(-cen[m]));
            if (!full) dv = mul2(dv, mask);
            const float keep = warp_sum_pair(add_(dv.x, dv.y), fma_(dv.x, dv.x, mul_(dv.y, dv.y)), lane);
            if (writer) wrow[m] = keep;
        };

        int pair = 0;
        for (int m = 1; m <= m_S1; ++m) { advance_pairs(t, pair, half, pair_tan); save_curve(m); }
        {   // ---- estimators at S1 ----
            PairState st; st.r1 = r1; st.r2 = r2; st.I1 = I1; st.I2 = I2;
            float2 x1, x2, c1, c2;
            zbc_payoffs(st, pl, K, x1, x2, c1, c2);
            const float2 tX = add2(x1, x2), tY = add2(c1, c2);
            const float2 tXX = fma2(x1, x1, mul2(x2, x2)), tYY = fma2(c1, c1, mul2(c2, c2));
            const float2 tXY = fma2(c1, x1, mul2(c2, x2));
            // pathwise vega of each twin (src/3:64-80): v = 1[P>K] * (-P B (xk B + t)) D - J D (P-K)^+
            auto vega_of = [&](float2 r, float2 I, float2 tg, float2 J) {
                const float2 z = mul2(mul2(r, splat(pl.negB)), splat(kLog2e));
                const float2 P = mul2(splat(pl.A), make_float2(mufu_ex2(z.x), mufu_ex2(z.y)));
                const float2 q = mul2(I, splat(-kLog2e));
                const float2 disc = make_float2(mufu_ex2(q.x), mufu_ex2(q.y));
                const float2 inner = fma2(splat(pl.xk), splat(pl.B), tg);
                float2 term1 = mul2(disc, mul2(mul2(P, splat(pl.negB)), inner));
                if (!(P.x > K)) term1.x = 0.0f;
                if (!(P.y > K)) term1.y = 0.0f;
                const float2 gk = add2(P, splat(-K));
                const float2 payoff = make_float2(fmaxf(0.0f, gk.x), fmaxf(0.0f, gk.y));
                const float2 zz = mul2(disc, J);
                return fma2(payoff, make_float2(-zz.x, -zz.y), term1);
            };
            const float2 v1 = vega_of(r1, I1, t1, J1), v2 = vega_of(r2, I2, t2, J2);
            const double mA = t.validA ? 1.0 : 0.0, mB = t.validB ? 1.0 : 0.0;
            const double vsA = (double)v1.x + (double)v2.x, vsB = (double)v1.y + (double)v2.y;
            const double ext[kFusedExtra] = {
                (double)tX.x * mA + (double)tX.y * mB, (double)tY.x * mA + (double)tY.y * mB,
                (double)tXX.x * mA + (double)tXX.y * mB, (double)tYY.x * mA + (double)tYY.y * mB,
                (double)tXY.x * mA + (double)tXY.y * mB, vsA * mA + vsB * mB, vsA * vsA * mA + vsB * vsB * mB,
                (double)v1.x * mA + (double)v1.y * mB};
#pragma unroll
            for (int k = 0; k < kFusedExtra; ++k) {
                const double w = warp_sum(ext[k]);
                if (lane == 0) wext[warp][k] = w;
            }
            if (FD) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    float2 y1, y2, k1, k2;
                    zbc_payoffs(b ? bp : bm, plans[1 + b], K, y1, y2, k1, k2);
                    const float2 mom5[5] = {add2(y1, y2), add2(k1, k2), fma2(y1, y1, mul2(y2, y2)),
                                            fma2(k1, k1, mul2(k2, k2)), fma2(k1, y1, mul2(k2, y2))};
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const double w = warp_sum((double)mom5[k].x * mA + (double)mom5[k].y * mB);
                        if (lane == 0) wext[warp][kFusedExtra + 5 * b + k] = w;
                    }
                }
            }
        }
        for (int m = m_S1 + 1; m < n_mat; ++m) { advance_pairs(t, pair, half, pair_plain); save_curve(m); }
        __syncthreads();
        for (int k = tid; k < 2 * n_mat; k += kThreads) {
            double acc = (double)wflt[k];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) acc += (double)wflt[w * 2 * n_mat + k];
            bacc[k] += acc;
        }
        if (tid < kExt) {
            double acc = wext[0][tid];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) acc += wext[w][tid];
            bacc[2 * n_mat + tid] += acc;
        }
    }
    __syncthreads();
    double* out = partials + (size_t)blockIdx.x * nq;
    for (int k = tid; k < nq; k += kThreads) out[k] = bacc[k];
}

}  // namespace hw1f
