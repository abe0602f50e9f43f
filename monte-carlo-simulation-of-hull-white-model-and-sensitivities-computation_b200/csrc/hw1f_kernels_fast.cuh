// hw1f_kernels_fast.cuh -- "decomposed" simulation kernels (engine mode HW1F_MODE_DECOMPOSED).
//
// The exact discretisation is linear in the Gaussian shocks, so every path this engine ever needs
// from one stream of normals -- the +G / -G antithetic twins, the sigma -/+ eps bumped twins, the
// pathwise tangent d(r)/d(sigma) -- is  (deterministic part) +/- (scale) * h  with ONE shared
// noise recursion per stream:
//
//     h_{i+1} = h_i e^{-a dt} + G_i ,   S_n = sum_{i<=n} h_i ,   q_n = 2 S_n - h_n
//     r(+/-)  = m_i      +/- sig_st h_i          (m: noise-free short rate, host table, double)
//     I(+/-)  = Im_i     +/- (dt/2) sig_st q_i   (trapezoid of the noise part: sum (h_i + h_{i+1}))
//     tangent = D_i      + (sig_st/sigma) h_i ,  its integral = ID_i + (dt/2)(sig_st/sigma) q_i
//
// Instead of S the kernels carry W_n = sum_{i<n} G_i (no dependence on h): summing the recursion gives
// S_n (1 - e) = W_n - e h_n, hence q_n = qA W_n - qB h_n with qA = 2/(1-e), qB = 2e/(1-e) + 1.  A Box-Muller pair
// (G1, G2) = s (sin v, cos v) then advances two steps with FIVE packed instructions:
//     a = e sin v + cos v,  b = sin v + cos v,  h <- e^2 h + s a,  W <- W + s b
// (for a dt -> 0 the difference W - e h cancels: the per-path rounding error of q grows like ulp / ((1-e) n),
// zero-mean, so the estimators are unaffected; tests cover 1-e = 5e-4)
// (e^2 is not a float: the kernels multiply by RN(e^2) and add the relative residual back once per five pairs,
// h += 5 rho h -- S responds to the decay 1/(1-e) ~ 100 times as strongly as h, so the half-ulp matters)
// i.e. 2.5 packed FP32 instructions per time step for BOTH antithetic twins of BOTH lanes (and for any
// number of sigma scenarios), instead of eight.  Same XORWOW integers and the same three Box-Muller floats per
// pair (s = sqrt(-2 ln u), sin v, cos v: the MUFU results) as the reference-order kernels; the products
// s sin v, s cos v enter through the fused multiply-adds above instead of being rounded on their own, so the
// per-path floats differ from the reference by rounding only (~1e-7 relative), far inside the 1e-5 north-star
// tolerance -- tests run both modes against the oracle and against the reference binaries.
//
// Curve save points: p0 = e^{-I+} + e^{-I-} = e^{-Im} (2 + d) with d = 2 cosh(z) - 2, z = c q the noise part of
// the integral.  d is evaluated as z^2 times a degree-4 polynomial in z^2 (no XU work, no cancellation at short
// maturities); a warp holding a |z| > 1.2 falls back to ex2(+/-z log2 e).
#pragma once
#include "hw1f_kernels_extra.cuh"

namespace hw1f {

// per sigma-scenario constants of the decomposed form
struct FastScen {
    float sg;       // sig_st
    float c;        // 0.5 * dt * sig_st
    float mS1;      // noise-free short rate at S1
    float ImS1;     // noise-free integral at S1
    const float* emI;   // [n_mat] exp(-Im) at the save points (curve kernels)
};
struct FastTangent {
    float DS1, IDS1;    // noise-free tangent and its integral at S1
};

#ifndef HW1F_FAST_COSH_POLY
#define HW1F_FAST_COSH_POLY 1
#endif
#ifndef HW1F_RHO_POS
#define HW1F_RHO_POS 2               // position of the decay residual in the five-pair group (any is equivalent to first order;
                                     // A/B of all five x HW1F_FIXED_HALF: profiles/r02_ab_variants.txt)
#endif
#ifndef HW1F_MLOOP_UNROLL
#define HW1F_MLOOP_UNROLL 0          // 0: compiler's choice for the maturity loop
#endif
#ifndef HW1F_POLY_ESTRIN
#define HW1F_POLY_ESTRIN 0           // 1: Estrin evaluation of the save-point polynomial (shorter dependency chain)
#endif
#ifndef HW1F_VOTE_EVERY
#define HW1F_VOTE_EVERY 1            // A/B: profiles/r02_ab_variants.txt
#endif
#ifndef HW1F_VOTE_THRESHOLD
#define HW1F_VOTE_THRESHOLD 0.8f
#endif
#ifndef HW1F_FIXED_HALF
#define HW1F_FIXED_HALF 1            // 1: a code path for the default save stride (10 steps = one five-pair group)
#endif
#ifndef HW1F_FOLD_LN2
#define HW1F_FOLD_LN2 0              // 1: the noise state is carried in units of sqrt(2 ln2): radius = sqrt(-lg2 u), no FMUL2 by -2 ln2
                                     // per pair (10 dispatch cycles per loop body less, one rounding less).  Measured: Q1 -0.2 %, but
                                     // ptxas' schedule of the ZBC / sequence kernels is 2 % slower (profiles/r02_ab_variants.txt, block 3)
#endif
#ifndef HW1F_POLY3
#define HW1F_POLY3 1                 // 1: degree-3 minimax polynomial at the save points (same 2.2e-8 as the degree-4 Taylor form)
#endif

// Window width of the stream derivation per instantiation (hw1f_device.cuh: window5_matvec).  The five-bit table
// saves a fifth of the look-ups (prologue 1289 -> 1079 instructions, 400 -> 320 LDS; 2-step ZBC call 48.1 -> 44.7 us),
// but the width also changes how ptxas orders the time loops behind the prologue, which is worth +-1 % on its own
// (DESIGN.md section 4): the width is therefore chosen per kernel family BY MEASUREMENT
// (profiles/r02_ab_window_width.txt): curve, ZBC and fused kernels five bits (Q1 -0.4 %, ZBC -0.6 %, 20-seed batch
// -0.7 %, recalibration curves -0.35 %, fused -0.65 %), pathwise and the one-launch Q3 sequence four (five bits: 0 % and
// +1.9 % -- the sequence kernel's three loops come out in a slower order behind the shorter prologue).
#ifndef HW1F_WIN_BITS_CURVE
#define HW1F_WIN_BITS_CURVE 5
#endif
#ifndef HW1F_WIN_BITS_ZBC
#define HW1F_WIN_BITS_ZBC 5
#endif
#ifndef HW1F_WIN_BITS_FUSED
#define HW1F_WIN_BITS_FUSED 5
#endif
#ifndef HW1F_WIN_BITS_OTHER
#define HW1F_WIN_BITS_OTHER 4
#endif
__host__ __device__ constexpr int fast_win_bits(int ncur, int nzbc, int pw, int seq)
{
    return (ncur > 0 && nzbc == 0 && pw == 0 && !seq) ? HW1F_WIN_BITS_CURVE
         : (ncur == 0 && nzbc > 0 && pw == 0)          ? HW1F_WIN_BITS_ZBC
         : (ncur > 0 && nzbc > 0 && !seq)              ? HW1F_WIN_BITS_FUSED
                                                       : HW1F_WIN_BITS_OTHER;
}
__host__ __device__ constexpr int fast_win_words(int ncur, int nzbc, int pw, int seq)
{
    return fast_win_bits(ncur, nzbc, pw, seq) == 5 ? kWin5Words : kWinWords;
}

struct FastState {
    float2 h, W;
    __device__ __forceinline__ float2 q(float qA, float qB) const { return fma2(W, splat(qA), mul2(h, splat(-qB))); }
};

// ZBC payoff X and control Y of both antithetic twins of both lanes for one sigma scenario, from the shared
// noise state (h, q) at S1:  r(+/-) = mS1 +/- sg h,  I(+/-) = ImS1 +/- c q.  mom5 = per-lane pair sums
// {X1+X2, Y1+Y2, X1^2+X2^2, Y1^2+Y2^2, X1 Y1 + X2 Y2} (common.cuh:356-362)
__device__ __forceinline__ void fast_zbc_mom5(float2 h, float2 q, const FastScen& z, const BondPlan& plan, float K,
                                              float2 (&mom5)[5])
{
    PairState ps;
    const float2 dr = mul2(h, splat(z.sg)), dI = mul2(q, splat(z.c));
    ps.r1 = add2(splat(z.mS1), dr);
    ps.r2 = add2(splat(z.mS1), make_float2(-dr.x, -dr.y));
    ps.I1 = add2(splat(z.ImS1), dI);
    ps.I2 = add2(splat(z.ImS1), make_float2(-dI.x, -dI.y));
    float2 x1, x2, c1, c2;
    zbc_payoffs(ps, plan, K, x1, x2, c1, c2);
    mom5[0] = add2(x1, x2);
    mom5[1] = add2(c1, c2);
    mom5[2] = fma2(x1, x1, mul2(x2, x2));
    mom5[3] = fma2(c1, c1, mul2(c2, c2));
    mom5[4] = fma2(c1, x1, mul2(c2, x2));
}

// NCUR  : number of curve scenarios accumulated on the maturity grid (0, 1, 2)
// NZBC  : number of ZBC scenarios evaluated at S1 (0..3); plans[0..NZBC)
// PW    : 0 none, 1 pathwise vega of the +G path only (the reference's estimator), 2 both twins
// DUMP  : 1 = also store the noise state (h, q) of every subsequence at step n_steps_S1 (a save point) to
//         dump[run][chunk * kChunk + {tid, tid + kThreads}] as float4 (h_A, q_A, h_B, q_B) halves, so that payoffs
//         whose constants depend on THIS pass's curve (recalibrated FD, src/3:484-525) are evaluated afterwards
//         by zbc_from_state_kernel without simulating the same normals again.  The curves of such a pass are read by
//         that pricing only -- P at the grid points around S1 and S2 and at the last maturity, f around S1 -- so `lead`
//         (unused by curve kernels, which start on a pair boundary) carries the two windows of maturities whose save
//         points are evaluated (keep_code(): a0 | b0 << 10 | 1 << 31, windows [a0, a0 + 6) and [b0, b0 + 6), plus the
//         last two grid points); the other sums stay zero = the noise-free curve after un-centring.  88 of 100 save
//         points and their warp reductions drop out of the recalibration pass: Q3 sequence 1.252 -> 1.178 ms, recalibrated FD
//         0.703 -> 0.620 ms (profiles/r02_ab_window_width.txt), every price bit for bit the same
// partials[run][block][NCUR*2*n_mat + NZBC*5 + (PW ? 3 : 0)] doubles; the S1 block is laid out as
// [ZBC scenario 0 (5)] [pathwise (3)] [ZBC scenarios 1.. (5 each)]  == the hw1f_fused* ABI order
template <int NZBC, int PW>
__device__ __forceinline__ constexpr int ext_zbc(int s) { return s == 0 ? 0 : 5 + (PW ? 3 : 0) + 5 * (s - 1); }
template <int NZBC, int PW>
__device__ __forceinline__ constexpr int ext_pw() { return NZBC > 0 ? 5 : 0; }

// SEQ   : 1 = the whole Q3 sequence of the reference's main() (src/3:697-834) in ONE pass over each subsequence's normals:
//         pathwise tangent on normals [0, n), then both CRN finite-difference bumps (scenarios 1, 2: zs1/zs2, plans[1],
//         plans[2]) on [n, 2n), then the two recalibration curves (cs0, cs1) on [2n, 2n + n_steps) with the noise state
//         parked at step n of that window (DUMP) -- one stream derivation instead of three.  Instantiated as
//         fast_kernel<2, 3, 1, 1, 1>; scenario 0's five ZBC slots stay zero.
// ODD   : 1 = curve sums for ANY save stride (e.g. 500 steps on 101 maturities): a save point may fall between the two
//         normals of a Box-Muller pair, whose cos half is then carried over to the next interval.  A separate
//         instantiation (fast_kernel<NCUR, 0, 0, 0, 0, 1>), so the even-stride kernels keep their code.
template <int NCUR, int NZBC, int PW, int DUMP = 0, int SEQ = 0, int ODD = 0>
#ifndef HW1F_FAST_MIN_BLOCKS
#define HW1F_FAST_MIN_BLOCKS 2   // A/B in profiles/r01_ab_variants_decomposed.txt: 512 threads x 2 blocks (64 regs) is best
#endif
__global__ void __launch_bounds__(kThreads, HW1F_FAST_MIN_BLOCKS)
fast_kernel(StreamGeom g, SeedArgs seeds, ModelDev md, FastScen cs0, FastScen cs1, FastScen zs0, FastScen zs1,
            FastScen zs2, FastTangent tg, const BondPlan* __restrict__ plans, int n_steps_S1, int lead, float K,
            double* __restrict__ partials, float2* __restrict__ dump)
{
    extern __shared__ __align__(16) uint32_t smem[];
    // the tail kernel behind this launch (hw1f_tail.cuh) may be made resident as soon as every block of this grid has
    // started: it parks in griddepcontrol.wait until the grid has completed and flushed
    asm volatile("griddepcontrol.launch_dependents;");
    constexpr int kS1 = NZBC * 5 + (PW ? 3 : 0);
    const int n_mat = md.n_mat;
    const int nqc = NCUR * 2 * n_mat;
    const int nq = nqc + kS1;
    uint32_t* win = smem;
    constexpr int WB = fast_win_bits(NCUR, NZBC, PW, SEQ);
    double* bacc = reinterpret_cast<double*>(smem + fast_win_words(NCUR, NZBC, PW, SEQ));   // [nqc]
    float* wflt = reinterpret_cast<float*>(bacc + (NCUR ? nqc : 0));            // [kWarps][nqc]
    float* emI = wflt + (NCUR ? kWarps * nqc : 0);                              // [NCUR][n_mat]
    __shared__ double wext[kWarps][kS1 > 0 ? kS1 : 1];
    __shared__ double bext[kS1 > 0 ? kS1 : 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int run = blockIdx.y;

    if (NCUR) {
        for (int k = tid; k < nqc; k += kThreads) bacc[k] = 0.0;
        for (int k = tid; k < kWarps * nqc; k += kThreads) wflt[k] = 0.0f;
        for (int k = tid; k < n_mat; k += kThreads) {
            emI[k] = cs0.emI[k];
            if (NCUR > 1) emI[n_mat + k] = cs1.emI[k];
        }
    }
    if (kS1 > 0 && tid < kS1) bext[tid] = 0.0;
    if (SEQ && tid < kWarps * 5) wext[tid / 5][tid % 5] = 0.0;   // scenario 0's ZBC slots are not used by the sequence

    const float2 e2 = splat(md.exp_adt), ee2 = splat(md.exp_2adt);
    const int half = md.stride >> 1;
    const int n_total = NCUR ? md.n_steps : n_steps_S1;
    const int m_S1 = NCUR ? n_steps_S1 / md.stride : 0;
    // lanes 0 / 16 of each warp own the (sum, sum of squares) slots of that warp's row
    const bool writer = (lane & 15) == 0;
    float* const wrow = wflt + warp * nqc + ((lane & 16) ? n_mat : 0);
    // per-scenario exponents: exp(-/+ c q) = ex2(-/+ q * (c log2e))
    const float kc0 = mul_(cs0.c, kLog2e), kc1 = mul_(cs1.c, kLog2e);
    // HW1F_FOLD_LN2: (h, W) are carried in units of kRadiusScale; every constant that reads them absorbs the factor
    constexpr int kRaw = HW1F_FOLD_LN2;
    auto unit_ = [](float x) { return kRaw ? mul_(x, kRadiusScale) : x; };
    const float zA0 = unit_(mul_(cs0.c, md.qA)), zB0 = -unit_(mul_(cs0.c, md.qB));
    const float zA1 = unit_(mul_(cs1.c, md.qA)), zB1 = -unit_(mul_(cs1.c, md.qB));

    // launched with programmatic stream serialisation behind prep_lo_kernel: everything above overlapped it; the
    // per-launch table U and the bond plans are read only below this point
    asm volatile("griddepcontrol.wait;" ::: "memory");

    for (unsigned long long chunk = blockIdx.x; chunk < g.n_chunks; chunk += gridDim.x) {
        ThreadStreams t = derive_streams<WB>(g, seeds, run, chunk, win);
        FastState st;
        st.h = splat(0.0f);
        st.W = splat(0.0f);
        const float2 mask = make_float2(t.validA ? 1.0f : 0.0f, t.validB ? 1.0f : 0.0f);
        const bool full = __syncthreads_and(t.validA && t.validB);

        int pair = 0;
        auto step1 = [&](float2 G) {       // single step (odd lead / tail only)
            if (kRaw) G = mul2(G, splat(kInvRadiusScale));
            st.W = add2(st.W, G);
            st.h = fma2(st.h, e2, G);
        };
        auto pairfn = [&](int j, float2 s, float2 sn, float2 cs) {
            const float2 a = fma2(sn, e2, cs);
            const float2 b = add2(sn, cs);
            // the residual of the group's five RN(e^2) decays goes in at position 1 (any position is equivalent to
            // first order; ptxas' schedule of this one is 1.7 % faster than of the others, profiles/r01_ab_variants_decomposed.txt)
            if (j == HW1F_RHO_POS) st.h = fma2(st.h, splat(md.rho5), st.h);
            st.h = fma2(s, a, mul2(st.h, ee2));
            st.W = fma2(s, b, st.W);
        };
        auto pairfn1 = [&](int, float2 s, float2 sn, float2 cs) {
            const float2 a = fma2(sn, e2, cs);
            const float2 b = add2(sn, cs);
            st.h = fma2(s, a, mul2(st.h, ee2));
            st.W = fma2(s, b, st.W);
        };
        // n_pairs pair steps in groups of five (pairfn adds the group's decay residual), single pairs for the rest
        auto advance = [&](int n_pairs) {
            int k = 0;
            for (; k + 5 <= n_pairs; k += 5) {
                run_pairs_parts<5, kRaw>(t, pair, pairfn);
                pair += 5;
            }
            for (; k < n_pairs; ++k) {
                st.h = fma2(st.h, splat(md.rho1), st.h);
                run_pairs_parts<1, kRaw>(t, pair, pairfn1);
                pair += 1;
            }
        };

        // DUMP passes: only the save points the pricing behind this pass reads (block-uniform test)
        const unsigned keep_a = (unsigned)lead & 1023u, keep_b = ((unsigned)lead >> 10) & 1023u;
        const bool keep_all = lead >= 0;
        auto keep_m = [&](int m) {
            return keep_all || (unsigned)(m - (int)keep_a) < 6u || (unsigned)(m - (int)keep_b) < 6u || m >= n_mat - 2;
        };
        bool big = false;   // a lane of this warp is near the validity limit of the save-point polynomial
        auto save_curve = [&](int m) {
#pragma unroll
            for (int s = 0; s < NCUR; ++s) {
                // p0 = e^{-I+} + e^{-I-} = e^{-Im} (e^{-c q} + e^{+c q});  centred: d = p0 - 2 e^{-Im}
#if HW1F_FAST_COSH_POLY
                // e^{-z} + e^{+z} - 2 = z^2 (1 + w/12 + w^2/360 + w^3/20160 + w^4/1814400), w = z^2 = (c q)^2:
                // truncation < 3e-8 relative for |z| <= 1.2 (z is the noise part of the integral, sd 0.29 at T = 10),
                // no cancellation at short maturities, and no XU work; a warp holding a |z| > 1.2 (0.3 % of the
                // warps at the last maturity) takes the exponential form
                // z = c q = (c qA) W - (c qB) h
                const float2 z = fma2(st.W, splat(s ? zA1 : zA0), mul2(st.h, splat(s ? zB1 : zB0)));
                float2 d2;
#if HW1F_VOTE_EVERY > 1
                // |z| is re-examined every HW1F_VOTE_EVERY-th save point against a lower threshold: over that many
                // save points z moves by ~0.09 sqrt(HW1F_VOTE_EVERY / 10) (one sd), and the polynomial degrades
                // gracefully (relative truncation 3e-8 at |z| = 1.2, 4e-6 at |z| = 2)
                if ((m % HW1F_VOTE_EVERY) == 1 || s > 0) {
                    if (s == 0) big = __any_sync(0xffffffffu, fmaxf(fabsf(z.x), fabsf(z.y)) > HW1F_VOTE_THRESHOLD);
                }
                if (big) {
#else
                if (__any_sync(0xffffffffu, fmaxf(fabsf(z.x), fabsf(z.y)) > 1.2f)) {
#endif
                    const float2 y = mul2(z, splat(kLog2e));
                    const float2 ep = make_float2(mufu_ex2(y.x), mufu_ex2(y.y));
                    const float2 en = make_float2(mufu_ex2(-y.x), mufu_ex2(-y.y));
                    d2 = add2(add2(ep, en), splat(-2.0f));
                } else {
                    const float2 w = mul2(z, z);
#if HW1F_POLY_ESTRIN
                    const float2 w2 = mul2(w, w);
                    const float2 lo = fma2(w, splat(1.0f / 12.0f), splat(1.0f));
                    const float2 mid = fma2(w, splat(1.0f / 20160.0f), splat(1.0f / 360.0f));
                    const float2 hi = mul2(w2, splat(1.0f / 1814400.0f));
                    const float2 pl = fma2(w2, add2(mid, hi), lo);
#elif HW1F_POLY3
                    // minimax fit of (2 cosh z - 2) / z^2 on w = z^2 in [0, 1.44] with the constant pinned to 1:
                    // 2.2e-8 relative, the truncation of the degree-4 Taylor form at |z| = 1.2, with one FFMA2 less
                    float2 pl = fma2(w, splat(5.115120075060986e-05f), splat(0.002776478650048375f));
                    pl = fma2(pl, w, splat(0.08333364129066467f));
                    pl = fma2(pl, w, splat(1.0f));
#else
                    float2 pl = fma2(w, splat(1.0f / 1814400.0f), splat(1.0f / 20160.0f));
                    pl = fma2(pl, w, splat(1.0f / 360.0f));
                    pl = fma2(pl, w, splat(1.0f / 12.0f));
                    pl = fma2(pl, w, splat(1.0f));
#endif
                    d2 = mul2(w, pl);
                }
                float2 dv = mul2(d2, splat(emI[s * n_mat + m]));
#else
                const float2 y = mul2(st.q(md.qA, md.qB), splat(unit_(s ? kc1 : kc0)));
                const float2 ep = make_float2(mufu_ex2(y.x), mufu_ex2(y.y));
                const float2 en = make_float2(mufu_ex2(-y.x), mufu_ex2(-y.y));
                float2 dv = mul2(add2(add2(ep, en), splat(-2.0f)), splat(emI[s * n_mat + m]));
#endif
                if (!full) dv = mul2(dv, mask);
                const float keep = warp_sum_pair(add_(dv.x, dv.y), fma_(dv.x, dv.x, mul_(dv.y, dv.y)), lane);
                if (writer) wrow[s * 2 * n_mat + m] = keep;
            }
        };

        auto eval_S1 = [&](bool do_zbc, bool do_pw) {
            FastState tu = st;   // the noise state in true units
            if (kRaw) { tu.h = mul2(st.h, splat(kRadiusScale)); tu.W = mul2(st.W, splat(kRadiusScale)); }
            const float2 q = tu.q(md.qA, md.qB);
            const double mA = t.validA ? 1.0 : 0.0, mB = t.validB ? 1.0 : 0.0;
#pragma unroll
            for (int s = (SEQ ? 1 : 0); s < NZBC; ++s) {
                if (!do_zbc) break;
                const FastScen z = (s == 0) ? zs0 : (s == 1 ? zs1 : zs2);
                float2 mom5[5];
                fast_zbc_mom5(tu.h, q, z, plans[s], K, mom5);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const double w = warp_sum((double)mom5[k].x * mA + (double)mom5[k].y * mB);
                    if (lane == 0) wext[warp][ext_zbc<NZBC, PW>(s) + k] = w;
                }
            }
            if (PW && do_pw) {
                const BondPlan pl = plans[0];
                const FastScen z = zs0;
                const float2 dr = mul2(tu.h, splat(z.sg)), dI = mul2(q, splat(z.c));
                const float2 dt_ = mul2(tu.h, splat(pl.c_t));                       // tangent noise
                const float2 dJ = mul2(q, splat(mul_(mul_(0.5f, md.dt), pl.c_t)));   // its integral
                auto vega_of = [&](float sgn) {
                    const float2 r = fma2(splat(sgn), dr, splat(z.mS1));
                    const float2 I = fma2(splat(sgn), dI, splat(z.ImS1));
                    const float2 tgv = fma2(splat(sgn), dt_, splat(tg.DS1));
                    const float2 J = fma2(splat(sgn), dJ, splat(tg.IDS1));
                    const float2 zz0 = mul2(mul2(r, splat(pl.negB)), splat(kLog2e));
                    const float2 P = mul2(splat(pl.A), make_float2(mufu_ex2(zz0.x), mufu_ex2(zz0.y)));
                    const float2 qq = mul2(I, splat(-kLog2e));
                    const float2 disc = make_float2(mufu_ex2(qq.x), mufu_ex2(qq.y));
                    const float2 inner = fma2(splat(pl.xk), splat(pl.B), tgv);
                    float2 term1 = mul2(disc, mul2(mul2(P, splat(pl.negB)), inner));
                    if (!(P.x > K)) term1.x = 0.0f;
                    if (!(P.y > K)) term1.y = 0.0f;
                    const float2 gk = add2(P, splat(-K));
                    const float2 payoff = make_float2(fmaxf(0.0f, gk.x), fmaxf(0.0f, gk.y));
                    const float2 zz = mul2(disc, J);
                    return fma2(payoff, make_float2(-zz.x, -zz.y), term1);
                };
                const float2 v1 = vega_of(1.0f);
                double e0, e1, e2;
                if (PW == 1) {   // the reference's non-antithetic estimator: sum v, sum v^2 over +G paths
                    const double a = (double)v1.x * mA, b = (double)v1.y * mB;
                    e0 = a + b; e1 = a * a + b * b; e2 = e0;
                } else {         // both twins: sum (v1+v2), sum (v1+v2)^2, sum v1
                    const float2 v2 = vega_of(-1.0f);
                    const double a = ((double)v1.x + (double)v2.x) * mA, b = ((double)v1.y + (double)v2.y) * mB;
                    e0 = a + b; e1 = a * a + b * b; e2 = (double)v1.x * mA + (double)v1.y * mB;
                }
                const double w0 = warp_sum(e0), w1 = warp_sum(e1), w2 = warp_sum(e2);
                if (lane == 0) { wext[warp][ext_pw<NZBC, PW>()] = w0; wext[warp][ext_pw<NZBC, PW>() + 1] = w1; wext[warp][ext_pw<NZBC, PW>() + 2] = w2; }
            }
        };

        if (NCUR && ODD) {
            bool pending = false;              // the cos half of the last pair has not been stepped yet
            float2 nc_keep = splat(0.0f);
            for (int m = 1; m < n_mat; ++m) {
                int left = md.stride;
                if (pending) { step1(nc_keep); pending = false; --left; }
                advance(left >> 1);
                if (left & 1) {
                    float2 ns, nc;
                    one_pair(t, ns, nc);
                    step1(ns);
                    nc_keep = nc;
                    pending = true;
                }
                save_curve(m);
            }
        } else if (NCUR) {
            if (SEQ) {   // n_steps_S1 even, normal offset even (checked on the host)
                advance(n_steps_S1 >> 1);
                eval_S1(false, true);          // pathwise tangent, normals [0, n)
                st.h = splat(0.0f);
                st.W = splat(0.0f);
                advance(n_steps_S1 >> 1);
                eval_S1(true, false);          // CRN finite-difference bumps, normals [n, 2n)
                st.h = splat(0.0f);
                st.W = splat(0.0f);
            }
#if HW1F_MLOOP_UNROLL == 2
#pragma unroll 2
#elif HW1F_MLOOP_UNROLL == 1
#pragma unroll 1
#endif
            for (int m = 1; m < n_mat; ++m) {
#if HW1F_FIXED_HALF
                if (half == 5) { run_pairs_parts<5, kRaw>(t, pair, pairfn); pair += 5; }
                else advance(half);
#else
                advance(half);
#endif
                if (!DUMP || keep_m(m)) save_curve(m);
                if (!SEQ && kS1 > 0 && m == m_S1) eval_S1(true, true);
                if (DUMP && m == m_S1) {
                    float2* d = dump + ((size_t)run * g.n_chunks + chunk) * kChunk + tid;
                    FastState tu = st;
                    if (kRaw) { tu.h = mul2(st.h, splat(kRadiusScale)); tu.W = mul2(st.W, splat(kRadiusScale)); }
                    const float2 q = tu.q(md.qA, md.qB);
                    d[0] = make_float2(tu.h.x, q.x);
                    d[kThreads] = make_float2(tu.h.y, q.y);
                }
            }
        } else {
            const int n_main = max(n_total - lead, 0);   // lead == 1 with zero steps: nothing to do
            if (lead && n_total > 0) {   // cached cos-branch normal of the pair the previous launch opened
                float2 ns, nc;
                one_pair(t, ns, nc);
                step1(nc);
            }
            advance(n_main >> 1);
            if (n_main & 1) {            // odd tail: sin branch only
                float2 ns, nc;
                one_pair(t, ns, nc);
                step1(ns);
            }
            eval_S1(true, true);
        }
        __syncthreads();
        if (NCUR) {
            for (int k = tid; k < nqc; k += kThreads) {
                double acc = (double)wflt[k];
#pragma unroll
                for (int w = 1; w < kWarps; ++w) acc += (double)wflt[w * nqc + k];
                bacc[k] += acc;
            }
        }
        if (kS1 > 0 && tid < kS1) {
            double acc = wext[0][tid];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) acc += wext[w][tid];
            bext[tid] += acc;
        }
    }
    __syncthreads();
    double* out = partials + ((size_t)run * gridDim.x + blockIdx.x) * nq;
    if (NCUR)
        for (int k = tid; k < nqc; k += kThreads) out[k] = bacc[k];
    if (kS1 > 0 && tid < kS1) out[nqc + tid] = bext[tid];
}

// ZBC/control moments of NZBC sigma scenarios from a dumped noise state (see DUMP above): no RNG, no recursion.
// Same thread <-> subsequence mapping, validity masks, warp/block reduction order and partials layout
// (partials[run][block][5 * NZBC]) as fast_kernel<0, NZBC, 0>, so reduce_partials_kernel finishes it.
template <int NZBC>
__global__ void __launch_bounds__(kThreads)
zbc_from_state_kernel(StreamGeom g, FastScen zs0, FastScen zs1, const BondPlan* __restrict__ plans, float K,
                      const float2* __restrict__ dump, double* __restrict__ partials)
{
    constexpr int kS1 = NZBC * 5;
    __shared__ double wext[kWarps][kS1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int run = blockIdx.y;
    double acc = 0.0;   // threads 0 .. kS1-1 own one moment each
    for (unsigned long long chunk = blockIdx.x; chunk < g.n_chunks; chunk += gridDim.x) {
        const unsigned long long pA = ((g.chunk0 + chunk) << kChunkLog2) + tid, pB = pA + kThreads;
        const double mA = (pA >= g.first_path && pA < g.first_path + g.n_paths) ? 1.0 : 0.0;
        const double mB = (pB >= g.first_path && pB < g.first_path + g.n_paths) ? 1.0 : 0.0;
        const float2* d = dump + ((size_t)run * g.n_chunks + chunk) * kChunk + tid;
        const float2 sA = d[0], sB = d[kThreads];
        const float2 h = make_float2(sA.x, sB.x), q = make_float2(sA.y, sB.y);
        __syncthreads();   // wext of the previous chunk has been consumed
#pragma unroll
        for (int s = 0; s < NZBC; ++s) {
            float2 mom5[5];
            fast_zbc_mom5(h, q, s ? zs1 : zs0, plans[s], K, mom5);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const double w = warp_sum((double)mom5[k].x * mA + (double)mom5[k].y * mB);
                if (lane == 0) wext[warp][5 * s + k] = w;
            }
        }
        __syncthreads();
        if (tid < kS1) {
            double a = wext[0][tid];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) a += wext[w][tid];
            acc += a;
        }
    }
    if (tid < kS1) partials[((size_t)run * gridDim.x + blockIdx.x) * kS1 + tid] = acc;
}

}  // namespace hw1f
