// hw1f_device.cuh -- device primitives of the B200 HW1F engine (sm_100a only).
//
// Numerics contract: every floating-point operation below is pinned with an explicit
// intrinsic or PTX instruction so that the per-path float sequence is exactly the one the
// reference kernels execute when built with `-O3 --use_fast_math` (SASS read off the
// reference's sm_100 build; see DESIGN.md "float sequence").  The packed f32x2 forms
// (FFMA2/FADD2/FMUL2, new on sm_100) round each lane like the scalar instruction, so packing
// two independent paths into one instruction does not change a single bit.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "hw1f kernels are written for sm_100a (B200) only"
#endif

namespace hw1f {

// Timing ablations (WRONG RESULTS, tools/ablation.sh only): HW1F_ABLATE_MUFU replaces every MUFU of the Box-Muller
// transform by one FMUL (same dispatch slots, no XU work); HW1F_ABLATE_RNG replaces the xorshift recurrence by one
// add per draw (the Weyl add, the conversions and the MUFUs stay).  They tell which resource a kernel waits for.
#ifndef HW1F_ABLATE_MUFU
#define HW1F_ABLATE_MUFU 0
#endif
#ifndef HW1F_ABLATE_RNG
#define HW1F_ABLATE_RNG 0
#endif
// ---- MUFU (XU pipe) wrappers: the approximations --use_fast_math selects ------------------
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#if HW1F_ABLATE_MUFU
__device__ __forceinline__ float mufu_lg2(float x) { return __fmul_rn(x, -0.999f); }
__device__ __forceinline__ float mufu_sqrt(float x) { return __fmul_rn(x, 0.998f); }
__device__ __forceinline__ float mufu_sqrt_abs(float x) { return __fmul_rn(x, 0.997f); }
__device__ __forceinline__ float mufu_sin(float x) { return __fmul_rn(x, 0.15f); }
__device__ __forceinline__ float mufu_cos(float x) { return __fmul_rn(x, 0.14f); }
#else
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// sqrt(|x|): the absolute value is an operand modifier of MUFU.SQRT (abs.f32 without .ftz folds; -x under -ftz=true is an FADD)
__device__ __forceinline__ float mufu_sqrt_abs(float x) { float y; asm("{.reg .f32 t; abs.f32 t, %1; sqrt.approx.ftz.f32 %0, t;}" : "=f"(y) : "f"(x)); return y; }
// sin/cos.approx expand to FMUL.RZ(x, 1/2pi) + MUFU.SIN/COS, exactly like __sincosf
__device__ __forceinline__ float mufu_sin(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_cos(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif

// ---- scalar FP32 with pinned rounding (never contracted by the compiler) --------------------
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }

// ---- packed FP32x2 (Blackwell FFMA2 / FADD2 / FMUL2) -----------------------------------------
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 splat(float x) { return make_float2(x, x); }

constexpr float kLog2e = 1.4426950216293334961f;   // expf(x)  -> ex2(x * kLog2e)
constexpr float kLn2 = 0.69314718246459960938f;    // logf(x)  -> lg2(x) * kLn2
constexpr float kNeg2Ln2 = -2.0f * kLn2;             // exact: power-of-two scaling of kLn2
constexpr uint32_t kWeyl = 362437u;                // curand_kernel.h:872

// ---- XORWOW ---------------------------------------------------------------------------------
#ifndef HW1F_SHR_IMAD
#define HW1F_SHR_IMAD 0
#endif
#if HW1F_SHR_IMAD
__constant__ uint32_t kTwo30 = 0x40000000u;
#endif
struct Xorwow {
    uint32_t v0, v1, v2, v3, v4;
    // the xorshift half of curand() (curand_kernel.h:866-871); the Weyl half is added by the caller
    __device__ __forceinline__ uint32_t next()
    {
#if HW1F_ABLATE_RNG
        v4 += v0;          // timing ablation only: one add instead of the six-instruction xorshift step
        return v4;
#endif
#if HW1F_SHR_IMAD
        // v0 >> 2 as the high word of v0 * 2^30: IMAD.HI on the FMA pipe instead of SHF on the (half-rate) ALU pipe, which
        // the five other integer instructions of a draw already load.  The factor sits in constant memory so that the
        // compiler cannot turn the multiplication back into a shift.
        const uint32_t t = v0 ^ __umulhi(v0, kTwo30);
#else
        const uint32_t t = v0 ^ (v0 >> 2);
#endif
        v0 = v1; v1 = v2; v2 = v3; v3 = v4;
        v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
        return v4;
    }
};

// u32 -> f32, round to nearest even (what I2FP.F32.U32 does)
__device__ __forceinline__ float u2f(uint32_t x)
{
#ifdef HW1F_I2F_ON_FMA
    // exact RN conversion without the conversion unit: two 16-bit halves are exact in FP32,
    // one FFMA merges them with a single rounding
    const float hi = __uint_as_float(0x4B000000u | (x >> 16)) - 8388608.0f;
    const float lo = __uint_as_float(0x4B000000u | (x & 0xffffu)) - 8388608.0f;
    return __fmaf_rn(hi, 65536.0f, lo);
#else
    return __uint2float_rn(x);
#endif
}

// _curand_box_muller (curand_normal.h:70-87) for two independent streams at once.
// x = first draw, y = second draw of each stream.  n_sin -> normal 2k, n_cos -> normal 2k+1.
// radius and direction separately: s = sqrt(-2 ln u), sn = sin v, cs = cos v (the three MUFU results per lane)
__device__ __forceinline__ void box_muller2_parts(uint32_t xa, uint32_t ya, uint32_t xb, uint32_t yb,
                                                  float2& s, float2& sn, float2& cs)
{
    const float2 fx = make_float2(u2f(xa), u2f(xb));
    const float2 fy = make_float2(u2f(ya), u2f(yb));
    const float2 u = fma2(fx, splat(__uint_as_float(0x2f800000u)), splat(__uint_as_float(0x2f000000u)));
    const float2 v = fma2(fy, splat(__uint_as_float(0x30c90fdbu)), splat(__uint_as_float(0x30490fdbu)));
    // reference: (lg2(u) * ln2) * -2  ->  one FMUL2 by (-2 ln2): scaling by -2 is exact, so
    // RN(RN(l*c) * -2) == RN(l * (-2c)) bit for bit (no subnormals can occur: |l*c| >= 5.9e-8 or 0)
    float2 l = make_float2(mufu_lg2(u.x), mufu_lg2(u.y));
    l = mul2(l, splat(kNeg2Ln2));
    s = make_float2(mufu_sqrt(l.x), mufu_sqrt(l.y));
    sn = make_float2(mufu_sin(v.x), mufu_sin(v.y));
    cs = make_float2(mufu_cos(v.x), mufu_cos(v.y));
}

// the same with the radius left unscaled: s = sqrt(-lg2 u) = sqrt(-2 ln u) / kRadiusScale.  The decomposed kernels carry
// their (linear) noise recursion in units of kRadiusScale and fold the factor into the constants that read it, which
// removes the FMUL2 by -2 ln2 from every pair (the negation is an operand modifier of MUFU.SQRT) and one rounding.
constexpr float kRadiusScale = 1.17741001f;      // sqrt(2 ln 2)
constexpr float kInvRadiusScale = 0.849321783f;  // 1 / sqrt(2 ln 2)
__device__ __forceinline__ void box_muller2_parts_raw(uint32_t xa, uint32_t ya, uint32_t xb, uint32_t yb,
                                                      float2& s, float2& sn, float2& cs)
{
    const float2 fx = make_float2(u2f(xa), u2f(xb));
    const float2 fy = make_float2(u2f(ya), u2f(yb));
    const float2 u = fma2(fx, splat(__uint_as_float(0x2f800000u)), splat(__uint_as_float(0x2f000000u)));
    const float2 v = fma2(fy, splat(__uint_as_float(0x30c90fdbu)), splat(__uint_as_float(0x30490fdbu)));
    s = make_float2(mufu_sqrt_abs(mufu_lg2(u.x)), mufu_sqrt_abs(mufu_lg2(u.y)));   // lg2 u <= 0
    sn = make_float2(mufu_sin(v.x), mufu_sin(v.y));
    cs = make_float2(mufu_cos(v.x), mufu_cos(v.y));
}

__device__ __forceinline__ void box_muller2(uint32_t xa, uint32_t ya, uint32_t xb, uint32_t yb,
                                            float2& n_sin, float2& n_cos)
{
    float2 s, sn, cs;
    box_muller2_parts(xa, ya, xb, yb, s, sn, cs);
    n_sin = mul2(s, sn);
    n_cos = mul2(s, cs);
}

// scalar version (single stream)
__device__ __forceinline__ void box_muller1(uint32_t x, uint32_t y, float& n_sin, float& n_cos)
{
    const float u = fma_(u2f(x), __uint_as_float(0x2f800000u), __uint_as_float(0x2f000000u));
    const float v = fma_(u2f(y), __uint_as_float(0x30c90fdbu), __uint_as_float(0x30490fdbu));
    const float s = mufu_sqrt(mul_(mul_(mufu_lg2(u), kLn2), -2.0f));
    n_sin = mul_(s, mufu_sin(v));
    n_cos = mul_(s, mufu_cos(v));
}

// evolve_hull_white_step (common.cuh:237-244) for two paths.  Reference sequence per path:
// FFMA (r'), FADD (r+r'), FMUL (*0.5), FFMA (*dt + I).  The halving is exact, so folding it into
// the constant (hdt = 0.5*dt, also exact) leaves the single rounding of the last FFMA unchanged:
// FFMA2, FADD2, FFMA2 -- three issue slots for two paths.
__device__ __forceinline__ void hw_step2(float2& r, float2& integral, float2 shock, float2 e2, float2 hdt2)
{
    const float2 rn = fma2(r, e2, shock);
    integral = fma2(add2(rn, r), hdt2, integral);
    r = rn;
}
__device__ __forceinline__ void hw_step1(float& r, float& integral, float shock, float e, float dt)
{
    const float rn = fma_(r, e, shock);
    const float h = mul_(add_(rn, r), 0.5f);
    integral = fma_(h, dt, integral);
    r = rn;
}

// ---- stateless stream derivation ---------------------------------------------------------------
// Window table of one 160x160 GF(2) matrix M: 40 nibble positions x 16 values x 5 words,
// entry[g][x] = M * (x << 4g).  20-byte entries: the 16 entries of a group sit in 16 distinct
// banks for every word (5 is odd), so a warp's lookups are conflict-free or broadcast.
constexpr int kWinGroups = 40;
constexpr int kWinWords = kWinGroups * 16 * 5;  // 3200 uint32 = 12800 B

__device__ __forceinline__ Xorwow window_matvec(const uint32_t* __restrict__ win, const uint32_t u[5])
{
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
#pragma unroll
    for (int w = 0; w < 5; ++w) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t idx = (u[w] >> (4 * j)) & 15u;
            const uint32_t* e = win + ((w * 8 + j) * 16 + idx) * 5;
            a0 ^= e[0]; a1 ^= e[1]; a2 ^= e[2]; a3 ^= e[3]; a4 ^= e[4];
        }
    }
    Xorwow s;
    s.v0 = a0; s.v1 = a1; s.v2 = a2; s.v3 = a3; s.v4 = a4;
    return s;
}

// The same product with FIVE-bit windows, the width the decomposed kernels use: 32 windows x 32 values, word k of
// entry[g][x] = (M * (x << 5g))[k] at win[(k * 32 + g) * 32 + x] (five planes).  The 32 entries of a window are
// consecutive words of one plane, so a warp's look-up (32 lanes picking among 32 entries) touches distinct banks or
// broadcasts: ONE shared-memory wavefront per LDS.32 whatever the lanes pick, and the five planes of a look-up sit at
// compile-time offsets from one address register (SHF + LOP3 + 5 LDS + 2.5 LOP3 per look-up).  The look-ups are the
// cost of the derivation -- the shared-memory pipe delivers one wavefront per clock and SM, and an LDS.64 / LDS.128
// whose lanes pick among entries costs 2 / 5 wavefronts (tools/probes/lds_probe.cu) -- so the window is as wide as a
// wavefront allows: 32 look-ups x 5 words per stream instead of 40 x 5.  A wider window would put two entries on one
// bank and pay every look-up twice.
constexpr int kWin5Groups = 32, kWin5Entries = 32, kWin5Plane = kWin5Groups * kWin5Entries;
constexpr int kWin5Words = 5 * kWin5Plane;   // 5120 uint32 = 20480 B
// global layout: the two tables of one matrix sit side by side, [four-bit table | five-bit table] per matrix
constexpr int kWinStride = kWinWords + kWin5Words;

__device__ __forceinline__ Xorwow window5_matvec(const uint32_t* __restrict__ win, const uint32_t u[5])
{
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
#pragma unroll
    for (int g = 0; g < kWin5Groups; ++g) {
        const int b = 5 * g, w = b >> 5, sh = b & 31;   // compile-time after unrolling
        uint32_t x;
        if (sh + 5 <= 32) x = u[w] >> sh;
        else x = __funnelshift_r(u[w], u[w + 1 < 5 ? w + 1 : 4], sh);   // a window across two state words (never the last)
        x &= 31u;
        const uint32_t* e = win + g * kWin5Entries + x;
        a0 ^= e[0]; a1 ^= e[kWin5Plane]; a2 ^= e[2 * kWin5Plane]; a3 ^= e[3 * kWin5Plane]; a4 ^= e[4 * kWin5Plane];
    }
    Xorwow s;
    s.v0 = a0; s.v1 = a1; s.v2 = a2; s.v3 = a3; s.v4 = a4;
    return s;
}

// one warp multiplies a 160-bit vector (replicated in every lane) by a row-image matrix in
// global memory: lane l owns bits l, l+32, .. of the input
__device__ __forceinline__ void warp_matvec(const uint32_t* __restrict__ mat, uint32_t v[5], int lane)
{
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
#pragma unroll
    for (int w = 0; w < 5; ++w) {
        const uint32_t m = 0u - ((v[w] >> lane) & 1u);
        const uint32_t* r = mat + (32 * w + lane) * 5;
        a0 ^= r[0] & m; a1 ^= r[1] & m; a2 ^= r[2] & m; a3 ^= r[3] & m; a4 ^= r[4] & m;
    }
    v[0] = __reduce_xor_sync(0xffffffffu, a0);
    v[1] = __reduce_xor_sync(0xffffffffu, a1);
    v[2] = __reduce_xor_sync(0xffffffffu, a2);
    v[3] = __reduce_xor_sync(0xffffffffu, a3);
    v[4] = __reduce_xor_sync(0xffffffffu, a4);
}

__device__ __forceinline__ float warp_sum(float x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = add_(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
// sums two values over the warp with 5 shuffles instead of 10: after the first exchange the
// lower half-warp carries partial sums of `a`, the upper half those of `b`.
// Returns the total of a in lanes 0..15 and the total of b in lanes 16..31 (fixed order).
__device__ __forceinline__ float warp_sum_pair(float a, float b, int lane)
{
    const bool up = (lane & 16) != 0;
    const float give = up ? a : b;
    float keep = up ? b : a;
    keep = add_(keep, __shfl_xor_sync(0xffffffffu, give, 16));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep = add_(keep, __shfl_xor_sync(0xffffffffu, keep, o));
    return keep;
}
__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

}  // namespace hw1f
