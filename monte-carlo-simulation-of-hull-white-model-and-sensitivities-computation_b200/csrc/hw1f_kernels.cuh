// hw1f_kernels.cuh -- the CUDA kernels of the HW1F engine (sm_100a).
//
// Thread mapping shared by every simulation kernel
//   * one block owns kChunk = 1024 consecutive RNG subsequences ("reference threads"),
//     aligned in ABSOLUTE path index, so all of them share the same high jump matrix;
//   * one thread owns two of them, A = base + tid and B = base + tid + 512, and keeps both
//     in the two lanes of packed FP32x2 registers (FFMA2/FADD2/FMUL2), which halves the
//     instruction count of the recursion (not its dispatch cycles: a packed instruction holds
//     the port for two);
//   * the XORWOW state of a subsequence is re-derived in the prologue:
//        v = J^(hi*L) * ( J^lo * T^offset * v0(seed) ),   path = hi*L + lo,
//     the bracket comes from the per-launch table U (prep_lo_kernel), J^(hi*L) is applied as
//     window look-ups into a table staged in shared memory: 32 five-bit windows (20.5 KB, decomposed curve / ZBC /
//     fused kernels) or 40 four-bit windows (12.8 KB, the others); hw1f_device.cuh.
//   * time loop in registers, drift table in shared memory as duplicated float2 (one LDS.128
//     feeds two steps of both lanes), Box-Muller phase static (no flag, no branch).
//   * reductions: warp shuffle tree -> shared -> one double partial row per block -> tail_kernel
//     (hw1f_tail.cuh: groups of blocks, then groups; fixed order, double).  No atomics anywhere:
//     results are bit-reproducible.
#pragma once
#include "hw1f_device.cuh"

namespace hw1f {

#ifndef HW1F_THREADS_LOG2
#define HW1F_THREADS_LOG2 9            // 512 threads per block, 2 blocks/SM (A/B: profiles/r01_ab_variants_decomposed.txt)
#endif
constexpr int kThreads = 1 << HW1F_THREADS_LOG2;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 2 * kThreads;   // subsequences per block
constexpr int kChunkLog2 = HW1F_THREADS_LOG2 + 1;
constexpr int kMaxRuns = 32;           // seed axis of the batched launches
constexpr int kMaxScen = 2;            // sigma scenarios sharing one set of normals
#ifndef HW1F_MIN_BLOCKS
#define HW1F_MIN_BLOCKS 2              // resident blocks per SM of the single-scenario kernels (64 regs at 512 threads)
#endif

struct SeedBlock {
    uint32_t v0[5];    // T^offset * v0(seed)  (offset jump folded in on the host)
    uint32_t d_start;  // Weyl word before the first draw of this launch
};
struct SeedArgs { SeedBlock s[kMaxRuns]; };

struct StreamGeom {
    const uint32_t* W;     // [n_hi][kWinStride] window tables of J^((hi_base+h)*L): four-bit windows (kWinWords), then the
                           // same matrix with five-bit windows (kWin5Words)
    const uint32_t* U;     // [n_runs][5][L]   lo vectors, SoA per run
    unsigned long long first_path, n_paths;
    unsigned long long chunk0;  // first_path / kChunk
    unsigned long long n_chunks; // chunks covering [first_path, first_path+n_paths)
    uint32_t hi_base;
    uint32_t L_log2;
};

struct ModelDev {
    float r0, exp_adt, dt, a;
    // decomposed kernels (hw1f_kernels_fast.cuh): exp_2adt = RN(exp_adt^2), the two-step decay as a float;
    // rho1 / rho5 = the relative residual (exp_adt^2 - exp_2adt)/exp_2adt of that rounding, once and five times
    // (applied as h += rho h so that the effective decay is exp_adt per step: S is 1/(1-e) times as sensitive to
    // the decay as h is); q = qA W - qB h with qA = 2/(1-e), qB = 2e/(1-e) + 1, e = exp_adt
    float exp_2adt, rho1, rho5, qA, qB;
    float inv_spacing, neg_spacing, spacing;
    int n_steps, n_mat, stride;
};

struct ScenDev {
    float sigma, sig_st;
    const float2* drift2;    // duplicated (d,d) drift table in global memory
    const float2* sdrift2;   // duplicated sensitivity drift table (pathwise only)
    const float* center;     // [n_mat] centring constants c_m of the curve accumulation (Q1 only)
    int slot;                // host bookkeeping: which of the engine's drift-table slots drift2 is
};

// path-independent pieces of P(S1,S2) = A exp(-B r) and of the pathwise tangent, computed ON
// THE DEVICE with the same MUFU sequence the reference executes per thread
// (common.cuh:180-225, src/3:15-19)
struct BondPlan {
    float A, B, negB, om2, xk, c_t;
};

// =================================================================================================
// per-launch preparation
// =================================================================================================
// interpolate() of common.cuh:187-196 as compiled: T/spacing -> T*10, alpha via one FFMA.  d0 = data[idx],
// d1 = data[idx+1]; idx >= n_mat-1 returns d0 (= data[n_mat-1])
__device__ __forceinline__ float interp_pts(float d0, float d1, int idx, float T, const ModelDev& md)
{
    if (idx >= md.n_mat - 1) return d0;
    const float al = mul_(fma_(__int2float_rn(idx), md.neg_spacing, T), md.inv_spacing);
    const float om = sub_(1.0f, al);
    return fma_(d0, om, mul_(al, d1));
}
__device__ __forceinline__ float interp_mkt(const float* __restrict__ data, float T, const ModelDev& md)
{
    const int idx = __float2int_rz(mul_(T, md.inv_spacing));
    if (idx >= md.n_mat - 1) return data[md.n_mat - 1];
    return interp_pts(data[idx], data[idx + 1], idx, T, md);
}

// the six market values a bond plan reads, picked on the host (the grid index (int)(T * inv_spacing) is plain
// IEEE arithmetic); the interpolation itself and everything after it stays on the device
struct MktPts {
    float P_S2[2], P_S1[2], f_S1[2];   // data[idx], data[idx+1] around S2 (P) and S1 (P, f)
    int idx_S2, idx_S1;
};

// one bond plan (common.cuh:180-225, src/3:15-19) in the reference's MUFU sequence.  use_pts: the six market values
// were picked on the host (every scenario prices on the same curves); else device-resident curves
__device__ __forceinline__ BondPlan compute_plan(const ModelDev& md, float sigma, float sig_st, float S1, float S2,
                                                 const float* __restrict__ P_mkt, const float* __restrict__ f_mkt,
                                                 const MktPts& pts, int use_pts)
{
    const float a = md.a;
    const float B = mul_(sub_(1.0f, mufu_ex2(mul_(mul_(sub_(S2, S1), a), -kLog2e))), mufu_rcp(a));
    const float P0T = use_pts ? interp_pts(pts.P_S2[0], pts.P_S2[1], pts.idx_S2, S2, md) : interp_mkt(P_mkt, S2, md);
    const float P0t = use_pts ? interp_pts(pts.P_S1[0], pts.P_S1[1], pts.idx_S1, S1, md) : interp_mkt(P_mkt, S1, md);
    const float f0t = use_pts ? interp_pts(pts.f_S1[0], pts.f_S1[1], pts.idx_S1, S1, md) : interp_mkt(f_mkt, S1, md);
    const float om2 = sub_(1.0f, mufu_ex2(mul_(mul_(mul_(a, -2.0f), S1), kLog2e)));
    float t3 = mul_(mul_(mul_(sigma, sigma), mufu_rcp(mul_(a, 4.0f))), om2);
    t3 = mul_(t3, B);
    t3 = mul_(t3, B);
    const float E = mufu_ex2(mul_(fma_(f0t, B, -t3), kLog2e));
    const float ratio = mul_(mufu_rcp(P0t), P0T);
    BondPlan p;
    p.B = B;
    p.negB = -B;
    p.A = mul_(ratio, E);
    p.om2 = om2;
    p.xk = mul_(om2, mul_(mufu_rcp(add_(a, a)), sigma));   // sigma/(2a) (1-e^{-2aS1}), src/3:17-18
    p.c_t = mul_(mufu_rcp(sigma), sig_st);                  // d_sig_st / d_sigma,       src/3:60
    return p;
}

// up to three bond plans computed as a side job of another launch (prep_lo_kernel's extra block, or the tail of a
// curve kernel for the recalibrated curves): no launch of its own
struct PlanJob {
    int n_scen;                    // 0: no job
    float sigma[3], sig_st[3];
    float S1, S2;
    MktPts pts;                    // use_pts = 1: host-picked market values
    int use_pts;
    const float* P_mkt[2];         // use_pts = 0: device curves of scenario 0 / 1
    const float* f_mkt[2];
    BondPlan* plans;               // output, n_scen entries
};

__device__ __forceinline__ void run_plan_job(const ModelDev& md, const PlanJob& job, int s)
{
    if (s >= job.n_scen) return;
    const int c = (s < 2) ? s : 1;
    job.plans[s] = compute_plan(md, job.sigma[s], job.sig_st[s], job.S1, job.S2, job.P_mkt[c], job.f_mkt[c], job.pts,
                                job.use_pts);
}

// a plan job as a launch of its own (batched entry points price many launches on one set of plans)
__global__ void plan_job_kernel(ModelDev md, PlanJob job) { run_plan_job(md, job, threadIdx.x); }

// U[run][w][lo] = word w of J^lo * v0'(run), lo < L.  One warp per (run, lo); one extra block (the last) runs the
// launch's plan job.  The dependent simulation kernel is launched with programmatic stream serialisation: this kernel
// releases it at once (its prologue overlaps the table build) and the dependent waits with griddepcontrol.wait
// before it touches U or the plans.
__global__ void __launch_bounds__(256)
prep_lo_kernel(SeedArgs seeds, int n_runs, uint32_t L_log2, const uint32_t* __restrict__ Jnib,
               uint32_t* __restrict__ U, ModelDev md, PlanJob job)
{
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31;
    const uint32_t L = 1u << L_log2;
    if (blockIdx.x == gridDim.x - 1) {          // plan block
        run_plan_job(md, job, threadIdx.x);
        return;
    }
    const unsigned long long wid = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (unsigned long long)n_runs * L) return;
    const int run = (int)(wid >> L_log2);
    const uint32_t lo = (uint32_t)(wid & (L - 1));
    uint32_t v[5];
#pragma unroll
    for (int w = 0; w < 5; ++w) v[w] = seeds.s[run].v0[w];
    // J^lo as the product of at most four nibble powers J^(n 16^j) (L <= 2^16): a dependent chain of four table
    // look-ups instead of one per set bit of lo
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t n = (lo >> (4 * j)) & 15u;
        if (n) warp_matvec(Jnib + (size_t)(j * 16 + n) * 800, v, lane);
    }
    if (lane < 5) {
        uint32_t out = v[0];
        if (lane == 1) out = v[1];
        if (lane == 2) out = v[2];
        if (lane == 3) out = v[3];
        if (lane == 4) out = v[4];
        U[((size_t)run * 5 + lane) * L + lo] = out;
    }
}

// W[h] = window table of J^((hi_base+h) * 2^L_log2).  One block of 160 threads per h.
__global__ void __launch_bounds__(160)
build_hi_kernel(uint32_t hi_base, uint32_t L_log2, const uint32_t* __restrict__ Jpow2, uint32_t* __restrict__ W)
{
    __shared__ uint32_t M[2][160][5];
    const int r = threadIdx.x;
    const uint32_t h = hi_base + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 5; ++k) M[0][r][k] = (k == (r >> 5)) ? (1u << (r & 31)) : 0u;
    __syncthreads();
    int cur = 0;
    for (uint32_t bit = 0; bit < 32; ++bit) {
        if (!((h >> bit) & 1u)) continue;   // block-uniform
        const uint32_t* J = Jpow2 + (size_t)(L_log2 + bit) * 800;
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
        for (int w = 0; w < 5; ++w) {
            uint32_t bits = M[cur][r][w];
            while (bits) {
                const int j = __ffs(bits) - 1;
                bits &= bits - 1;
                const uint32_t* row = J + (32 * w + j) * 5;
                a0 ^= row[0]; a1 ^= row[1]; a2 ^= row[2]; a3 ^= row[3]; a4 ^= row[4];
            }
        }
        M[cur ^ 1][r][0] = a0; M[cur ^ 1][r][1] = a1; M[cur ^ 1][r][2] = a2;
        M[cur ^ 1][r][3] = a3; M[cur ^ 1][r][4] = a4;
        __syncthreads();
        cur ^= 1;
    }
    uint32_t* out = W + (size_t)blockIdx.x * kWinStride;
    for (int e = r; e < kWinGroups * 16; e += 160) {
        const int g = e >> 4, x = e & 15;
        uint32_t a[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < 4; ++b)
            if ((x >> b) & 1) {
#pragma unroll
                for (int k = 0; k < 5; ++k) a[k] ^= M[cur][4 * g + b][k];
            }
#pragma unroll
        for (int k = 0; k < 5; ++k) out[e * 5 + k] = a[k];
    }
    // the same matrix as a five-bit window table (layout: window5_matvec, hw1f_device.cuh)
    uint32_t* out5 = out + kWinWords;
    for (int e = r; e < kWin5Plane; e += 160) {
        const int g = e >> 5, x = e & 31;
        uint32_t a[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < 5; ++b)
            if ((x >> b) & 1) {
#pragma unroll
                for (int k = 0; k < 5; ++k) a[k] ^= M[cur][5 * g + b][k];
            }
#pragma unroll
        for (int k = 0; k < 5; ++k) out5[k * kWin5Plane + e] = a[k];
    }
}

// =================================================================================================
// common prologue: stage the window table, derive the two XORWOW states of this thread
// =================================================================================================
struct ThreadStreams {
    Xorwow A, B;
    bool validA, validB;
    uint32_t dcur, dcurB;
};

// WB = window width of the table staged in `win`: 4 (kWinWords words, reference-order kernels) or 5 (kWin5Words)
template <int WB = 4>
__device__ __forceinline__ ThreadStreams derive_streams(const StreamGeom& g, const SeedArgs& seeds, int run,
                                                         unsigned long long chunk,
                                                         uint32_t* win /* smem, kWinWords or kWin5Words */)
{
    constexpr int kWords = (WB == 5) ? kWin5Words : kWinWords;
    const int tid = threadIdx.x;
    const unsigned long long base = (g.chunk0 + chunk) << kChunkLog2;
    __syncthreads();   // every warp is done with the previous chunk's window table
    const uint32_t L = 1u << g.L_log2;
    const uint32_t hi = (uint32_t)(base >> g.L_log2) - g.hi_base;
    const uint4* wsrc = reinterpret_cast<const uint4*>(g.W + (size_t)hi * kWinStride + (WB == 5 ? kWinWords : 0));
    for (int i = tid; i < kWords / 4; i += kThreads) reinterpret_cast<uint4*>(win)[i] = wsrc[i];
    const uint32_t loA = (uint32_t)(base & (L - 1)) + tid;
    const uint32_t loB = loA + kThreads;
    const uint32_t* Urun = g.U + (size_t)run * 5 * L;
    uint32_t uA[5], uB[5];
#pragma unroll
    for (int w = 0; w < 5; ++w) {
        uA[w] = Urun[(size_t)w * L + loA];
        uB[w] = Urun[(size_t)w * L + loB];
    }
    __syncthreads();
    ThreadStreams t;
    t.A = (WB == 5) ? window5_matvec(win, uA) : window_matvec(win, uA);
    t.B = (WB == 5) ? window5_matvec(win, uB) : window_matvec(win, uB);
    const unsigned long long pA = base + tid, pB = pA + kThreads;
    t.validA = (pA >= g.first_path) && (pA < g.first_path + g.n_paths);
    t.validB = (pB >= g.first_path) && (pB < g.first_path + g.n_paths);
    // keep the Weyl word in a vector register: xorshift + Weyl + immediate is then ONE IADD3 per
    // draw instead of a uniform-datapath add plus a vector add
    // (bit 63 of a path index is always 0; the dependence on pA / pB only defeats uniform-register
    // allocation and common-subexpression merging: with one private Weyl register per stream,
    // xorshift + Weyl + immediate is a single IADD3 per draw)
    t.dcur = seeds.s[run].d_start + (uint32_t)(pA >> 63);
    t.dcurB = seeds.s[run].d_start + (uint32_t)(pB >> 63);
    return t;
}

// NP Box-Muller pairs = 2*NP normals per stream; Weyl words are compile-time offsets of dcur.
// f(pair_index, n_sin, n_cos) consumes the two normals of a pair (two time steps).
template <int NP, class PairFn>
__device__ __forceinline__ void run_pairs(ThreadStreams& t, int pair, PairFn&& f)
{
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const uint32_t xa = t.A.next() + (t.dcur + kWeyl * (2 * j + 1));
        const uint32_t xb = t.B.next() + (t.dcurB + kWeyl * (2 * j + 1));
        const uint32_t ya = t.A.next() + (t.dcur + kWeyl * (2 * j + 2));
        const uint32_t yb = t.B.next() + (t.dcurB + kWeyl * (2 * j + 2));
        float2 ns, nc;
        box_muller2(xa, ya, xb, yb, ns, nc);
        f(pair + j, ns, nc);
    }
    t.dcur += kWeyl * (2 * NP);
    t.dcurB += kWeyl * (2 * NP);
}

// advance `n_pairs` Box-Muller pairs starting at pair index `pair` (unrolled by 5 pairs = 10 draws,
// the period after which the XORWOW register rotation is the identity)
template <class PairFn>
__device__ __forceinline__ void advance_pairs(ThreadStreams& t, int& pair, int n_pairs, PairFn&& f)
{
    int k = 0;
    for (; k + 5 <= n_pairs; k += 5) { run_pairs<5>(t, pair, f); pair += 5; }
    for (; k < n_pairs; ++k) { run_pairs<1>(t, pair, f); pair += 1; }
}

// the same walk handing out radius and direction of each pair separately: f(pair_index, s, sin v, cos v)
// (decomposed kernels fold the products s sin v, s cos v into their fused multiply-adds); j = position in the group
template <int NP, int RAW = 0, class PartsFn>
__device__ __forceinline__ void run_pairs_parts(ThreadStreams& t, int pair, PartsFn&& f)
{
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const uint32_t xa = t.A.next() + (t.dcur + kWeyl * (2 * j + 1));
        const uint32_t xb = t.B.next() + (t.dcurB + kWeyl * (2 * j + 1));
        const uint32_t ya = t.A.next() + (t.dcur + kWeyl * (2 * j + 2));
        const uint32_t yb = t.B.next() + (t.dcurB + kWeyl * (2 * j + 2));
        float2 s, sn, cs;
        if (RAW) box_muller2_parts_raw(xa, ya, xb, yb, s, sn, cs);   // radius in units of kRadiusScale
        else box_muller2_parts(xa, ya, xb, yb, s, sn, cs);
        f(j, s, sn, cs);
    }
    t.dcur += kWeyl * (2 * NP);
    t.dcurB += kWeyl * (2 * NP);
}
// one isolated pair (lead / tail handling of odd normal offsets and odd step counts)
__device__ __forceinline__ void one_pair(ThreadStreams& t, float2& ns, float2& nc)
{
    const uint32_t xa = t.A.next() + (t.dcur + kWeyl), xb = t.B.next() + (t.dcurB + kWeyl);
    const uint32_t ya = t.A.next() + (t.dcur + 2 * kWeyl), yb = t.B.next() + (t.dcurB + 2 * kWeyl);
    t.dcur += 2 * kWeyl;
    t.dcurB += 2 * kWeyl;
    box_muller2(xa, ya, xb, yb, ns, nc);
}

// =================================================================================================
// Q1: antithetic bond curve (simulate_zcb, market_data.cuh:25-79)
// =================================================================================================
// partials[run][block][nq] doubles, nq = NSCEN * 2 * n_mat: per scenario sum_m d, sum_m d^2 with
// d = p0_m - c_m.  Centring (c_m = the noise-free value, host-computed) makes the float shuffle
// trees lose nothing: the variance of p0 at short maturities is 1e-10 of its square and would
// vanish in float32 otherwise.  tail_kernel (hw1f_tail.cuh) undoes the centring in double.
// Blocks stride over chunks; per-warp float trees -> shared floats -> double block accumulators.
// ODD = 1: any save stride (a save point may fall between the two normals of a Box-Muller pair); separate instantiation
template <int NSCEN, int ODD = 0>
__global__ void __launch_bounds__(kThreads, (NSCEN > 1 ? 1 : HW1F_MIN_BLOCKS))
bond_curve_kernel(StreamGeom g, SeedArgs seeds, ModelDev md, ScenDev sc0, ScenDev sc1, double* __restrict__ partials)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int n_steps = md.n_steps, n_mat = md.n_mat;
    const int nq1 = 2 * n_mat;   // per scenario
    const int nq = nq1 * NSCEN;
    const int n_pairs_tot = (n_steps + 1) >> 1;   // an odd step count leaves the last slot half used (tables are zero padded)
    uint32_t* win = smem;
    float4* drift4 = reinterpret_cast<float4*>(smem + kWinWords);                   // [NSCEN][n_steps/2] (d_i,d_i,d_i+1,d_i+1)
    double* bacc = reinterpret_cast<double*>(drift4 + (size_t)NSCEN * n_pairs_tot);  // [nq] block accumulators
    float* wflt = reinterpret_cast<float*>(bacc + nq);                              // [kWarps][nq] this chunk
    float* cen = wflt + kWarps * nq;                                                // [NSCEN][n_mat]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int run = blockIdx.y;

    for (int i = tid; i < n_pairs_tot; i += kThreads) {
        drift4[i] = reinterpret_cast<const float4*>(sc0.drift2)[i];
        if (NSCEN > 1) drift4[n_pairs_tot + i] = reinterpret_cast<const float4*>(sc1.drift2)[i];
    }
    for (int k = tid; k < nq; k += kThreads) bacc[k] = 0.0;
    for (int k = tid; k < kWarps * nq; k += kThreads) wflt[k] = 0.0f;
    for (int k = tid; k < n_mat; k += kThreads) {
        cen[k] = sc0.center[k];
        if (NSCEN > 1) cen[n_mat + k] = sc1.center[k];
    }

    float2 sgP[NSCEN], sgM[NSCEN];
#pragma unroll
    for (int s = 0; s < NSCEN; ++s) {
        const float sg = s ? sc1.sig_st : sc0.sig_st;
        sgP[s] = splat(sg);
        sgM[s] = splat(-sg);
    }
    const float2 e2 = splat(md.exp_adt), hdt2 = splat(mul_(0.5f, md.dt));
    const int half = md.stride >> 1;
    // lane 0 stores the sum, lane 16 the sum of squares (see warp_sum_pair)
    const bool writer = (lane & 15) == 0;
    float* const wrow = wflt + warp * nq + ((lane & 16) ? n_mat : 0);

    for (unsigned long long chunk = blockIdx.x; chunk < g.n_chunks; chunk += gridDim.x) {
        ThreadStreams t = derive_streams(g, seeds, run, chunk, win);   // contains the __syncthreads()
        float2 r1[NSCEN], r2[NSCEN], I1[NSCEN], I2[NSCEN];
#pragma unroll
        for (int s = 0; s < NSCEN; ++s) {
            r1[s] = r2[s] = splat(md.r0);
            I1[s] = I2[s] = splat(0.0f);
        }
        const float mA = t.validA ? 1.0f : 0.0f, mB = t.validB ? 1.0f : 0.0f;
        const bool full = __syncthreads_and(t.validA && t.validB);

        auto pairfn = [&](int pk, float2 ns, float2 nc) {
#pragma unroll
            for (int s = 0; s < NSCEN; ++s) {
                const float4 d = drift4[s * n_pairs_tot + pk];
                const float2 da = make_float2(d.x, d.y), db = make_float2(d.z, d.w);
                hw_step2(r1[s], I1[s], fma2(ns, sgP[s], da), e2, hdt2);
                hw_step2(r2[s], I2[s], fma2(ns, sgM[s], da), e2, hdt2);
                hw_step2(r1[s], I1[s], fma2(nc, sgP[s], db), e2, hdt2);
                hw_step2(r2[s], I2[s], fma2(nc, sgM[s], db), e2, hdt2);
            }
        };

        // one step of every scenario with normal G; which = 0 / 1: first / second step of pair pk
        auto stepfn = [&](int pk, int which, float2 G) {
#pragma unroll
            for (int s = 0; s < NSCEN; ++s) {
                const float4 d = drift4[s * n_pairs_tot + pk];
                const float2 dd = which ? make_float2(d.z, d.w) : make_float2(d.x, d.y);
                hw_step2(r1[s], I1[s], fma2(G, sgP[s], dd), e2, hdt2);
                hw_step2(r2[s], I2[s], fma2(G, sgM[s], dd), e2, hdt2);
            }
        };
        int pair = 0;
        bool pending = false;                  // ODD: the cos half of pair `pair - 1` has not been stepped yet
        float2 nc_keep = splat(0.0f);
        for (int m = 1; m < n_mat; ++m) {
            if (ODD) {
                int left = md.stride;
                if (pending) { stepfn(pair - 1, 1, nc_keep); pending = false; --left; }
                advance_pairs(t, pair, left >> 1, pairfn);
                if (left & 1) {
                    float2 ns, nc;
                    one_pair(t, ns, nc);
                    stepfn(pair, 0, ns);
                    ++pair;
                    nc_keep = nc;
                    pending = true;
                }
            } else {
                advance_pairs(t, pair, half, pairfn);
            }
#pragma unroll
            for (int s = 0; s < NSCEN; ++s) {
                // p0_m = expf(-integral1) + expf(-integral2)   (market_data.cuh:60)
                const float2 a = mul2(I1[s], splat(-kLog2e)), b = mul2(I2[s], splat(-kLog2e));
                const float2 p0 = add2(make_float2(mufu_ex2(a.x), mufu_ex2(a.y)),
                                       make_float2(mufu_ex2(b.x), mufu_ex2(b.y)));
                float2 dv = add2(p0, splat(-cen[s * n_mat + m]));
                if (!full) dv = mul2(dv, make_float2(mA, mB));   // block-uniform branch: ragged edge chunks only
                const float keep = warp_sum_pair(add_(dv.x, dv.y), fma_(dv.x, dv.x, mul_(dv.y, dv.y)), lane);
                if (writer) wrow[s * nq1 + m] = keep;
            }
        }
        __syncthreads();
        for (int k = tid; k < nq; k += kThreads) {
            double acc = (double)wflt[k];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) acc += (double)wflt[w * nq + k];
            bacc[k] += acc;
        }
    }
    __syncthreads();
    double* out = partials + ((size_t)run * gridDim.x + blockIdx.x) * nq;
    for (int k = tid; k < nq; k += kThreads) out[k] = bacc[k];
}

// =================================================================================================
// Q2b / FD bumps: antithetic ZBC payoff with control variate (simulate_ZBC_control_variate,
// common.cuh:286-409); NSCEN sigma scenarios share the normals (common random numbers)
// =================================================================================================
// handles any start parity / step count: `lead` = 1 when the launch starts on the cos half of a
// Box-Muller pair (odd normal offset), partials[run][block][NSCEN*5] doubles
template <int NSCEN>
__global__ void __launch_bounds__(kThreads, HW1F_MIN_BLOCKS)
zbc_kernel(StreamGeom g, SeedArgs seeds, ModelDev md, ScenDev sc0, ScenDev sc1, const BondPlan* __restrict__ plans,
           int n_steps_S1, int lead, float K, double* __restrict__ partials)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* win = smem;
    // drift of steps lead, lead+1, ..: pair k of the main loop reads (d,d,d',d') with one LDS.128
    const int n_main = max(n_steps_S1 - lead, 0);         // steps after the optional lead step
    const int n_slots = (n_main + 1) >> 1;                // float4 slots per scenario
    float4* drift4 = reinterpret_cast<float4*>(smem + kWinWords);   // [NSCEN][n_slots]
    __shared__ double wpart[kWarps][NSCEN * 5];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int run = blockIdx.y;
    for (int i = tid; i < n_slots; i += kThreads) {
#pragma unroll
        for (int s = 0; s < NSCEN; ++s) {
            const float2* src = s ? sc1.drift2 : sc0.drift2;
            const int i0 = lead + 2 * i, i1 = i0 + 1;
            const float2 a = src[i0];
            const float2 b = (i1 < n_steps_S1) ? src[i1] : make_float2(0.0f, 0.0f);
            drift4[s * n_slots + i] = make_float4(a.x, a.y, b.x, b.y);
        }
    }
    if (tid < kWarps * NSCEN * 5) (&wpart[0][0])[tid] = 0.0;
    float2 sgP[NSCEN], sgM[NSCEN];
#pragma unroll
    for (int s = 0; s < NSCEN; ++s) {
        const float sg = s ? sc1.sig_st : sc0.sig_st;
        sgP[s] = splat(sg);
        sgM[s] = splat(-sg);
    }
    const float2 e2 = splat(md.exp_adt), hdt2 = splat(mul_(0.5f, md.dt));

    for (unsigned long long chunk = blockIdx.x; chunk < g.n_chunks; chunk += gridDim.x) {
        ThreadStreams t = derive_streams(g, seeds, run, chunk, win);
        float2 r1[NSCEN], r2[NSCEN], I1[NSCEN], I2[NSCEN];
#pragma unroll
        for (int s = 0; s < NSCEN; ++s) {
            r1[s] = r2[s] = splat(md.r0);
            I1[s] = I2[s] = splat(0.0f);
        }
        auto step1 = [&](int s, float2 d, float2 G) {
            hw_step2(r1[s], I1[s], fma2(G, sgP[s], d), e2, hdt2);
            hw_step2(r2[s], I2[s], fma2(G, sgM[s], d), e2, hdt2);
        };
        auto pairfn = [&](int pk, float2 ns, float2 nc) {
#pragma unroll
            for (int s = 0; s < NSCEN; ++s) {
                const float4 d = drift4[s * n_slots + pk];
                step1(s, make_float2(d.x, d.y), ns);
                step1(s, make_float2(d.z, d.w), nc);
            }
        };

        if (lead && n_steps_S1 > 0) {   // cached cos-branch normal of the pair the previous launch opened
            float2 ns, nc;
            one_pair(t, ns, nc);
#pragma unroll
            for (int s = 0; s < NSCEN; ++s) step1(s, (s ? sc1.drift2 : sc0.drift2)[0], nc);
        }
        int pair = 0;
        advance_pairs(t, pair, n_main >> 1, pairfn);
        if (n_main & 1) {               // odd tail: sin branch only (the cos value stays cached)
            float2 ns, nc;
            one_pair(t, ns, nc);
#pragma unroll
            for (int s = 0; s < NSCEN; ++s) {
                const float4 d = drift4[s * n_slots + pair];
                step1(s, make_float2(d.x, d.y), ns);
            }
        }

        const double mA = t.validA ? 1.0 : 0.0, mB = t.validB ? 1.0 : 0.0;
#pragma unroll
        for (int s = 0; s < NSCEN; ++s) {
            const BondPlan pl = plans[s];
            const float2 z1 = mul2(mul2(r1[s], splat(pl.negB)), splat(kLog2e));
            const float2 z2 = mul2(mul2(r2[s], splat(pl.negB)), splat(kLog2e));
            const float2 P1 = mul2(splat(pl.A), make_float2(mufu_ex2(z1.x), mufu_ex2(z1.y)));
            const float2 P2 = mul2(splat(pl.A), make_float2(mufu_ex2(z2.x), mufu_ex2(z2.y)));
            const float2 q1 = mul2(I1[s], splat(-kLog2e)), q2 = mul2(I2[s], splat(-kLog2e));
            const float2 d1 = make_float2(mufu_ex2(q1.x), mufu_ex2(q1.y));
            const float2 d2 = make_float2(mufu_ex2(q2.x), mufu_ex2(q2.y));
            const float2 c1 = mul2(P1, d1), c2 = mul2(P2, d2);               // control = discount * P
            const float2 g1 = add2(P1, splat(-K)), g2 = add2(P2, splat(-K));
            const float2 x1 = mul2(d1, make_float2(fmaxf(0.0f, g1.x), fmaxf(0.0f, g1.y)));
            const float2 x2 = mul2(d2, make_float2(fmaxf(0.0f, g2.x), fmaxf(0.0f, g2.y)));
            const float2 tX = add2(x1, x2);                                  // common.cuh:356-362
            const float2 tY = add2(c1, c2);
            const float2 tXX = fma2(x1, x1, mul2(x2, x2));
            const float2 tYY = fma2(c1, c1, mul2(c2, c2));
            const float2 tXY = fma2(c1, x1, mul2(c2, x2));
            const float2 v[5] = {tX, tY, tXX, tYY, tXY};
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const double w = warp_sum((double)v[k].x * mA + (double)v[k].y * mB);
                if (lane == 0) wpart[warp][s * 5 + k] += w;
            }
        }
    }
    __syncthreads();
    if (tid < NSCEN * 5) {
        double acc = wpart[0][tid];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) acc += wpart[w][tid];
        partials[((size_t)run * gridDim.x + blockIdx.x) * (NSCEN * 5) + tid] = acc;
    }
}

// =================================================================================================
// Q3 pathwise vega (simulate_sensitivity, src/3:22-96): NOT antithetic; r and d(r)/d(sigma)
// driven by the same normal.  partials[run][block][2] doubles: sum v, sum v^2
// =================================================================================================
__global__ void __launch_bounds__(kThreads, HW1F_MIN_BLOCKS)
pathwise_kernel(StreamGeom g, SeedArgs seeds, ModelDev md, ScenDev sc, const BondPlan* __restrict__ plans,
                int n_steps_S1, int lead, float K, double* __restrict__ partials)
{
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* win = smem;
    float4* dd = reinterpret_cast<float4*>(smem + kWinWords);   // [n_steps_S1] (d,d,sd,sd)
    __shared__ double wpart[kWarps][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int run = blockIdx.y;
    for (int i = tid; i < n_steps_S1; i += kThreads) {
        const float2 a = sc.drift2[i], b = sc.sdrift2[i];
        dd[i] = make_float4(a.x, a.y, b.x, b.y);
    }
    if (tid < kWarps * 2) (&wpart[0][0])[tid] = 0.0;
    const BondPlan pl = plans[0];
    const float2 sg = splat(sc.sig_st), ct = splat(pl.c_t), e2 = splat(md.exp_adt), hdt2 = splat(mul_(0.5f, md.dt));
    const int n_main = max(n_steps_S1 - lead, 0);

    for (unsigned long long chunk = blockIdx.x; chunk < g.n_chunks; chunk += gridDim.x) {
        ThreadStreams t = derive_streams(g, seeds, run, chunk, win);
        float2 r = splat(md.r0), tg = splat(0.0f), Ir = splat(0.0f), It = splat(0.0f);
        auto step1 = [&](int i, float2 G) {
            const float4 d = dd[i];
            hw_step2(r, Ir, fma2(G, sg, make_float2(d.x, d.y)), e2, hdt2);
            hw_step2(tg, It, fma2(ct, G, make_float2(d.z, d.w)), e2, hdt2);
        };
        auto pairfn = [&](int pk, float2 ns, float2 nc) {
            step1(lead + 2 * pk, ns);
            step1(lead + 2 * pk + 1, nc);
        };
        if (lead && n_steps_S1 > 0) {
            float2 ns, nc;
            one_pair(t, ns, nc);
            step1(0, nc);
        }
        int pair = 0;
        advance_pairs(t, pair, n_main >> 1, pairfn);
        if (n_main & 1) {
            float2 ns, nc;
            one_pair(t, ns, nc);
            step1(lead + 2 * pair, ns);
        }

        // src/3:64-80 as compiled
        const float2 z = mul2(mul2(r, splat(pl.negB)), splat(kLog2e));
        const float2 P = mul2(splat(pl.A), make_float2(mufu_ex2(z.x), mufu_ex2(z.y)));
        const float2 q = mul2(Ir, splat(-kLog2e));
        const float2 disc = make_float2(mufu_ex2(q.x), mufu_ex2(q.y));
        const float2 inner = fma2(splat(pl.xk), splat(pl.B), tg);
        const float2 y = mul2(mul2(P, splat(pl.negB)), inner);
        float2 term1 = mul2(disc, y);
        if (!(P.x > K)) term1.x = 0.0f;
        if (!(P.y > K)) term1.y = 0.0f;
        const float2 gk = add2(P, splat(-K));
        const float2 payoff = make_float2(fmaxf(0.0f, gk.x), fmaxf(0.0f, gk.y));
        const float2 zz = mul2(disc, It);
        const float2 v = fma2(payoff, make_float2(-zz.x, -zz.y), term1);
        const double vA = t.validA ? (double)v.x : 0.0, vB = t.validB ? (double)v.y : 0.0;
        const double s1 = warp_sum(vA + vB);
        const double s2 = warp_sum(vA * vA + vB * vB);
        if (lane == 0) { wpart[warp][0] += s1; wpart[warp][1] += s2; }
    }
    __syncthreads();
    if (tid < 2) {
        double acc = wpart[0][tid];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) acc += wpart[w][tid];
        partials[((size_t)run * gridDim.x + blockIdx.x) * 2 + tid] = acc;
    }
}

// =================================================================================================
// plain block-sum kernel: moments[run][q] = sum over blocks (fixed order, double).  The simulation launches are
// finished by tail_kernel (hw1f_tail.cuh); this one remains for the reduction benchmark's deterministic method
// =================================================================================================
// partials[run][block][stride]; sums entries q = 0..gridDim.x-1 -> moments[run*out_stride + q]
template <class T>
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const T* __restrict__ partials, int n_blocks, int stride, double* __restrict__ moments,
                       int out_stride)
{
    __shared__ double sh[256];
    const int q = blockIdx.x, run = blockIdx.y, tid = threadIdx.x;
    const T* p = partials + (size_t)run * n_blocks * stride + q;
    double acc = 0.0;
    for (int b = tid; b < n_blocks; b += 256) acc += (double)p[(size_t)b * stride];
    sh[tid] = acc;
    __syncthreads();
#pragma unroll
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) sh[tid] += sh[tid + o];
        __syncthreads();
    }
    if (tid == 0) moments[(size_t)run * out_stride + q] = sh[0];
}

// =================================================================================================
// epilogues (path-independent; device-side so that rcp/lg2/ex2 are the same MUFU results the
// reference's epilogue kernels produce)
// =================================================================================================
// compute_average_and_forward (market_data.cuh:101-127) + standard error of P, executed by ONE block (any size);
// s_P: n_mat floats of shared scratch.  P_se may be null.  Used by curve_epilogue_kernel and by the tail of the
// simulation kernels (hw1f_tail.cuh).
__device__ __forceinline__ void curve_epilogue_block(const double* __restrict__ moments, int n_mat, unsigned long long n_pairs,
                                                     float inv_dT, float* __restrict__ P, float* __restrict__ f,
                                                     float* __restrict__ P_se, float* s_P)
{
    // (float)n_paths of the reference (an int there); same value from the 64-bit count, and defined beyond 2^31
    const float n_paths_f = __ull2float_rn(2ull * n_pairs);
    for (int m = threadIdx.x; m < n_mat; m += blockDim.x) {
        // P_sum[0] = 2.0f * N_PATHS (market_data.cuh:76-78)
        const float sum = (m == 0) ? mul_(2.0f, __ull2float_rn(n_pairs)) : __double2float_rn(moments[m]);
        const float avg = mul_(sum, mufu_rcp(n_paths_f));
        s_P[m] = avg;
        P[m] = avg;
        if (P_se) {
            // pair-average sample x = p0/2: var = (sum x^2 - (sum x)^2/n)/(n-1); se = sqrt(var/n)
            double se = 0.0;
            if (m > 0 && n_pairs > 1) {
                const double n = (double)n_pairs;
                const double sx = 0.5 * moments[m], sxx = 0.25 * moments[n_mat + m];
                double var = (sxx - sx * sx / n) / (n - 1.0);
                if (var < 0.0) var = 0.0;
                se = sqrt(var / n);
            }
            P_se[m] = (float)se;
        }
    }
    __syncthreads();
    for (int m = threadIdx.x; m < n_mat; m += blockDim.x) {
        const int first = (m == 0) ? 0 : m - 1;
        const int last = (m == n_mat - 1) ? n_mat - 1 : m + 1;
        const float nscale = ((m == 0) || (m == n_mat - 1)) ? -1.0f : -0.5f;
        const float c = mul_(nscale, inv_dT);
        const float lf = mul_(mufu_lg2(s_P[first]), kLn2);
        const float dl = fma_(mufu_lg2(s_P[last]), kLn2, -lf);
        f[m] = mul_(c, dl);
    }
    __syncthreads();
}

__global__ void curve_epilogue_kernel(const double* __restrict__ moments, int n_mat, unsigned long long n_pairs,
                                      float inv_dT, float* __restrict__ P, float* __restrict__ f,
                                      float* __restrict__ P_se)
{
    extern __shared__ float s_P[];
    curve_epilogue_block(moments, n_mat, n_pairs, inv_dT, P, f, P_se, s_P);
}

// recover_theta (src/2_option_pricing.cu:14-35) + compute_derivative (common.cuh:250-258)
__global__ void theta_kernel(const float* __restrict__ f, int n_mat, float a, float sigma, float spacing,
                             float th_a0, float th_b0, float th_a1, float th_b1, float th_break,
                             float* __restrict__ theta_rec, float* __restrict__ theta_ref, float* __restrict__ Ts)
{
    const float coef = mul_(mul_(sigma, sigma), mufu_rcp(add_(a, a)));
    const float m2a = mul_(a, -2.0f);
    for (int i = threadIdx.x; i < n_mat; i += blockDim.x) {
        const float T = mul_(__int2float_rn(i), spacing);
        float df;
        if (i == 0) df = mul_(sub_(f[1], f[0]), mufu_rcp(spacing));
        else if (i == n_mat - 1) df = mul_(sub_(f[i], f[i - 1]), mufu_rcp(spacing));
        else df = mul_(sub_(f[i + 1], f[i - 1]), mufu_rcp(add_(spacing, spacing)));
        const float e = mufu_ex2(mul_(mul_(m2a, T), kLog2e));
        const float om = sub_(1.0f, e);
        theta_rec[i] = fma_(coef, om, fma_(f[i], a, df));
        theta_ref[i] = (T < th_break) ? fma_(T, th_b0, th_a0) : fma_(T, th_b1, th_a1);
        Ts[i] = T;
    }
}

// (int)(S1 / d_dt) as the reference's fast-math build evaluates it (common.cuh:322)
__global__ void steps_probe_kernel(float S1, float dt, int* out)
{
    *out = __float2int_rz(mul_(mufu_rcp(dt), S1));
}

// =================================================================================================
// sample trajectories (simulate_paths_show, market_data.cuh:136-160) and RNG introspection
// =================================================================================================
// one thread per path, scalar; used for n_show ~ 32 paths and by the parity tests
__global__ void sample_paths_kernel(StreamGeom g, SeedArgs seeds, ModelDev md, ScenDev sc, int n_show, int lead,
                                    const float* __restrict__ drift, float* __restrict__ out,
                                    uint32_t* __restrict__ dbg_state, uint32_t* __restrict__ dbg_draws, int n_draws,
                                    float* __restrict__ dbg_normals, int n_normals)
{
    // scalar re-derivation straight from the global tables (slow path, a handful of threads)
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_show) return;
    const unsigned long long p = g.first_path + q;
    const uint32_t L = 1u << g.L_log2;
    const uint32_t hi = (uint32_t)(p >> g.L_log2) - g.hi_base;
    const uint32_t lo = (uint32_t)(p & (L - 1));
    uint32_t u[5];
#pragma unroll
    for (int w = 0; w < 5; ++w) u[w] = g.U[(size_t)w * L + lo];
    Xorwow st = window_matvec(g.W + (size_t)hi * kWinStride, u);
    uint32_t d = seeds.s[0].d_start;
    if (dbg_state) {
        uint32_t* o = dbg_state + (size_t)q * 6;
        o[0] = d; o[1] = st.v0; o[2] = st.v1; o[3] = st.v2; o[4] = st.v3; o[5] = st.v4;
    }
    if (dbg_draws) {
        Xorwow c = st;
        uint32_t dd = d;
        for (int k = 0; k < n_draws; ++k) { dd += kWeyl; dbg_draws[(size_t)q * n_draws + k] = c.next() + dd; }
    }
    // curand_normal stream with its one-value cache (curand_normal.h:313-326)
    bool have = false;
    float extra = 0.0f;
    auto next_normal = [&]() {
        if (!have) {
            d += kWeyl; const uint32_t x = st.next() + d;
            d += kWeyl; const uint32_t y = st.next() + d;
            float a, b;
            box_muller1(x, y, a, b);
            extra = b; have = true;
            return a;
        }
        have = false;
        return extra;
    };
    if (lead) (void)next_normal();   // position on the cached cos value
    if (dbg_normals) {
        for (int k = 0; k < n_normals; ++k) dbg_normals[(size_t)q * n_normals + k] = next_normal();
        return;
    }
    if (out) {
        float r = md.r0, I = 0.0f;
        float* o = out + (size_t)q * (md.n_steps + 1);
        o[0] = r;
        for (int i = 1; i <= md.n_steps; ++i) {
            const float G = next_normal();
            hw_step1(r, I, fma_(G, sc.sig_st, drift[i - 1]), md.exp_adt, md.dt);
            o[i] = r;
        }
    }
}

}  // namespace hw1f
