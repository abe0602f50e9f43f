// hw1f_probe.cuh -- pipe-throughput micro-kernels.  They give the roofline denominators for this
// instruction-bound path (FP32 FMA pipe, XU/MUFU pipe, ALU pipe, conversion unit, issue slots),
// measured on the same GPU and in the same process as the HW1F kernels.
#pragma once
#include "hw1f_device.cuh"

namespace hw1f {

constexpr int kProbeUnroll = 64;   // instructions of the probed kind per chain per loop trip
constexpr int kProbeChains = 8;    // independent dependency chains per thread

// which: 0 FFMA, 1 FFMA2, 2 MUFU.EX2, 3 LOP3, 4 I2FP.F32.U32 (+LOP3 feeding it), 5 MUFU.EX2+I2FP mix,
//        6 FFMA2 + LOP3 + MUFU mix in the HW1F ratio (13:10:4), 7 FMUL2,
//        8 MUFU.LG2, 9 MUFU.SQRT, 10 MUFU.SIN (+ its FMUL.RZ), 11 MUFU.COS (+ FMUL.RZ),
//        12 the Box-Muller MUFU mix LG2 : SQRT : SIN : COS = 1 : 1 : 1 : 1 alone,
//        13 that mix with the decomposed Q1 loop's other work per four MUFU (14 LOP3/SHF, 4 FFMA2, 2 I2FP; the two FMUL.RZ
//           come with sin/cos): what a perfectly interleaved instruction stream of the same pipe mix reaches
template <int WHICH>
__global__ void __launch_bounds__(256) probe_kernel(int iters, float seedf, uint32_t seedu, float* sink)
{
    float a[kProbeChains];
    float2 a2[kProbeChains];
    uint32_t u[kProbeChains];
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) {
        a[c] = seedf + (float)(threadIdx.x + c) * 1e-3f;
        a2[c] = make_float2(a[c], a[c] + 1.0f);
        u[c] = seedu ^ (threadIdx.x * 2654435761u + c);
    }
    const float m = 0.999f + seedf * 1e-9f, b = 1e-3f;
    const float2 m2 = splat(m), b2 = splat(b);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kProbeUnroll; ++k) {
#pragma unroll
            for (int c = 0; c < kProbeChains; ++c) {
                if (WHICH == 0) a[c] = fma_(a[c], m, b);
                if (WHICH == 1) a2[c] = fma2(a2[c], m2, b2);
                if (WHICH == 2) a[c] = mufu_ex2(a[c]);
                if (WHICH == 3) u[c] = (u[c] ^ (u[c] << 3)) ^ seedu;   // SHF/LOP3 chain
                if (WHICH == 4) { u[c] ^= __float_as_uint(a[c]); a[c] = __uint2float_rn(u[c]); }
                if (WHICH == 5) { a[c] = mufu_ex2(a[c]); a2[c].x = __uint2float_rn(u[c] ^ __float_as_uint(a2[c].x)); }
                if (WHICH == 7) a2[c] = mul2(a2[c], m2);
                if (WHICH == 8) a[c] = mufu_lg2(fabsf(a[c]));
                if (WHICH == 9) a[c] = mufu_sqrt_abs(a[c]);
                if (WHICH == 10) a[c] = mufu_sin(a[c]);
                if (WHICH == 11) a[c] = mufu_cos(a[c]);
                if (WHICH == 12) {
                    if ((c & 3) == 0) a[c] = mufu_lg2(fabsf(a[c]));
                    if ((c & 3) == 1) a[c] = mufu_sqrt_abs(a[c]);
                    if ((c & 3) == 2) a[c] = mufu_sin(a[c]);
                    if ((c & 3) == 3) a[c] = mufu_cos(a[c]);
                }
            }
            if (WHICH == 13) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    a[4 * g] = mufu_lg2(fabsf(a[4 * g]));
                    a[4 * g + 1] = mufu_sqrt_abs(a[4 * g + 1]);
                    a[4 * g + 2] = mufu_sin(a[4 * g + 2]);
                    a[4 * g + 3] = mufu_cos(a[4 * g + 3]);
#pragma unroll
                    for (int c = 0; c < 14; ++c) u[c & 7] = (u[c & 7] ^ (u[c & 7] << 3)) ^ seedu;
#pragma unroll
                    for (int c = 0; c < 4; ++c) a2[4 * g + c] = fma2(a2[4 * g + c], m2, b2);
                    a2[4 * g].x = __uint2float_rn(u[2 * g]);
                    a2[4 * g + 1].y = __uint2float_rn(u[2 * g + 1]);
                }
            }
            if (WHICH == 6) {
                // one "HW1F step pair" worth of pipe pressure: 13 FP2, 10 ALU, 4 MUFU, 2 I2FP
#pragma unroll
                for (int c = 0; c < 13; ++c) a2[c & 7] = fma2(a2[c & 7], m2, b2);
#pragma unroll
                for (int c = 0; c < 10; ++c) u[c & 7] = (u[c & 7] ^ (u[c & 7] << 3)) ^ seedu;
#pragma unroll
                for (int c = 0; c < 4; ++c) a[c] = mufu_ex2(a[c]);
                a[4] = __uint2float_rn(u[0]);
                a[5] = __uint2float_rn(u[1]);
            }
        }
    }
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) acc += a[c] + a2[c].x + a2[c].y + __uint_as_float(u[c] & 0x3fffffffu);
    if (acc == 123.456f) sink[0] = acc;   // never true; keeps the chains alive
}

}  // namespace hw1f
