/*
 * oracle/ref/probe_steps.cu -- TEST INFRASTRUCTURE.
 * Reads, on the actual GPU, the value of `(int)(S1 / d_dt)` exactly as the
 * reference computes it under --use_fast_math (include/common.cuh:322,
 * src/3_sensitivity_analysis.cu:46: MUFU.RCP(d_dt) * S1 then F2I.TRUNC), and a
 * few MUFU outputs used to sanity-check the oracle's libm stand-ins.
 */
#include <cstdio>
#include <cuda_runtime.h>

__constant__ float c_dt;

__global__ void probe(float S1, int* n, float* out)
{
    *n = (int)(S1 / c_dt);
    out[0] = 1.0f / c_dt;
    out[1] = S1 / c_dt;
    out[2] = __expf(-0.1f);
    out[3] = __logf(0.5f);
    out[4] = __sinf(1.0f);
    out[5] = __cosf(1.0f);
    out[6] = sqrtf(2.0f);
}

int main()
{
    const float dt = 10.0f / 1000;
    cudaMemcpyToSymbol(c_dt, &dt, sizeof(float));
    int* d_n; float* d_out;
    cudaMalloc(&d_n, sizeof(int));
    cudaMalloc(&d_out, 8 * sizeof(float));
    const float S[] = {5.0f, 1.0f, 2.5f, 10.0f};
    for (int i = 0; i < 4; i++) {
        probe<<<1, 1>>>(S[i], d_n, d_out);
        int n; float o[8];
        cudaMemcpy(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(o, d_out, 8 * sizeof(float), cudaMemcpyDeviceToHost);
        printf("{\"S1\": %.1f, \"n_steps\": %d, \"rcp_dt\": %.9g, \"S1_over_dt\": %.9g, \"expf_m0.1\": %.9g, \"logf_0.5\": %.9g, \"sinf_1\": %.9g, \"cosf_1\": %.9g, \"sqrtf_2\": %.9g}\n",
               S[i], n, o[0], o[1], o[2], o[3], o[4], o[5], o[6]);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
