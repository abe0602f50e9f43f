// Shared-memory wavefronts of table look-ups: how many cycles does a warp-wide LDS.32 / LDS.64 / LDS.128 take when the 32 lanes
// pick among 16 entries (the window look-up of the stream derivation), all the same entry, or 32 distinct ones?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_probe lds_probe.cu ; run on the GPU box (one SM is enough).
// READ IT UNDER ncu, not by its own clock: the compiler moves the loop-invariant loads out of the timing loop, so the printed
// cycles are not per load.  The valid figure is wavefronts per executed load instruction,
//   ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum ./lds_probe
// Measured on B200 (profiles/r02_lds_probe_ncu.csv): LDS.32 among 16 entries 1.00 wavefront, LDS.64 among 16 entries 1.99
// (the two half-warps are served separately; identical addresses are not merged across them), LDS.128 among 16 entries 5.4
// (bank conflicts between entries 8 apart).  A 20-byte table entry therefore costs five wavefronts however it is loaded.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int BYTES, int MODE>   // MODE 0: among 16 entries (random per lane), 1: all lanes the same entry, 2: lane-distinct entries
__global__ void __launch_bounds__(1024) lds_kernel(int iters, unsigned* out, long long* cycles)
{
    __shared__ __align__(16) uint32_t tab[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = i * 2654435761u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t h = threadIdx.x * 747796405u + 2891336453u;
    uint32_t acc = 0;
    // byte offsets of 16 look-ups per lane, fixed over the loop (index arithmetic must not be what is measured)
    uint32_t off[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        h = h * 1664525u + 1013904223u;
        const uint32_t idx = (MODE == 0) ? ((h >> 24) & 15u) : (MODE == 1 ? (uint32_t)(k & 15) : (uint32_t)lane);
        off[k] = (k & 7u) * 2048u + idx * BYTES;
    }
    const char* tb = reinterpret_cast<const char*>(tab);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t rot = (it & 1u) * 16384u;   // a uniform, iteration-dependent offset: the loads cannot be hoisted
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(tb + off[k] + rot);
            if (BYTES == 4) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); acc ^= v; }
            if (BYTES == 8) { uint32_t v0, v1; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(a)); acc ^= v0 ^ v1; }
            if (BYTES == 16) { uint32_t v0, v1, v2, v3; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a)); acc ^= v0 ^ v1 ^ v2 ^ v3; }
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int BYTES, int MODE>
void run(const char* name)
{
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 1024 * sizeof(unsigned)); cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2048;
    lds_kernel<BYTES, MODE><<<1, 1024>>>(64, out, cyc);
    lds_kernel<BYTES, MODE><<<1, 1024>>>(iters, out, cyc);
    long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double warp_instr = 32.0 * iters * 16;   // 32 warps on the SM
    printf("%-44s %6.2f cycles per warp-wide load (%d B per lane)\n", name, (double)c / warp_instr, BYTES);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<4, 0>("LDS.32  lanes pick among 16 entries");
    run<8, 0>("LDS.64  lanes pick among 16 entries");
    run<16, 0>("LDS.128 lanes pick among 16 entries");
    run<4, 1>("LDS.32  all lanes the same entry");
    run<8, 1>("LDS.64  all lanes the same entry");
    run<16, 1>("LDS.128 all lanes the same entry");
    run<4, 2>("LDS.32  32 distinct consecutive entries");
    run<8, 2>("LDS.64  32 distinct consecutive entries");
    run<16, 2>("LDS.128 32 distinct consecutive entries");
    return cudaDeviceSynchronize() != cudaSuccess;
}
