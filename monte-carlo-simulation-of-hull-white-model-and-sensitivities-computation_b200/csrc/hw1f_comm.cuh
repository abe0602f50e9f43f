// hw1f_comm.cuh -- device side of the path's single exchange step: an all-reduce of the moment vector
// over NVLink peer memory (CUDA IPC mailboxes), shared by the stand-alone kernel (hw1f_comm.cu) and by
// the tail of the simulation kernels (hw1f_tail.cuh), where the LAST block of the reduction posts the
// vector to the peers itself -- no separate launch between reduction and collective.
#pragma once
#include <cuda_runtime.h>

namespace hw1f {

constexpr int kCommMaxWorld = 8;
constexpr int kCommMaxCount = 512;             // doubles per collective (Q1: 202, fused: 220, recalibration curves: 404)
constexpr unsigned kCommSpinLimit = 20000000u;  // a few seconds of polling: far beyond any healthy skew, still finite

struct Mailbox {
    double slots[2][kCommMaxWorld][kCommMaxCount];   // [epoch parity][source rank][value]
    unsigned flags[2][kCommMaxWorld];
    unsigned timeouts;
};

struct CommDev {
    Mailbox* peer[kCommMaxWorld];   // peer[r] = rank r's mailbox as mapped into this process
    int rank, world;                // world <= 1: no exchange
};

// system-scope release / acquire on the mailbox flags (the payload stores above the release are plain stores)
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// In-place SUM all-reduce of data[0, count) by ONE block (every thread of the block calls this; count <=
// kCommMaxCount; blockDim.x >= world):
//   1. post this rank's vector into its slot of EVERY rank's mailbox (plain stores over NVLink / NVSwitch), fence at
//      system scope, raise the per-slot flag with a release store;
//   2. wait (bounded spin) until every rank's flag for this epoch has arrived in the LOCAL mailbox;
//   3. sum the slots in RANK ORDER: bit-identical on every rank and from run to run.
// Mailboxes are double-buffered by epoch parity: a rank reaches epoch e+2 only after every peer has posted e+1, i.e.
// after every peer finished reading e.  A lost peer poisons the result with NaN (every *_finish rejects it) and bumps
// the mailbox's time-out counter instead of hanging the GPU.
__device__ __forceinline__ void block_peer_allreduce(const CommDev& c, double* __restrict__ data, int count, unsigned epoch)
{
    __shared__ int s_timed_out;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int p = epoch & 1u;
    if (tid == 0) s_timed_out = 0;
    for (int i = tid; i < count; i += nthr) {
        const double mine = data[i];
        for (int r = 0; r < c.world; ++r) c.peer[r]->slots[p][c.rank][i] = mine;
    }
    __threadfence_system();   // every thread's payload stores are ordered before the block-wide barrier ...
    __syncthreads();
    if (tid < c.world) st_release_sys(&c.peer[tid]->flags[p][c.rank], epoch);   // ... and published by the release
    Mailbox* me = c.peer[c.rank];
    if (tid < c.world) {
        unsigned spins = 0;
        while (ld_acquire_sys(&me->flags[p][tid]) != epoch) {
            if (++spins > kCommSpinLimit) { atomicAdd(&me->timeouts, 1u); s_timed_out = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    const bool bad = s_timed_out != 0;
    for (int i = tid; i < count; i += nthr) {
        double acc = 0.0;
        for (int r = 0; r < c.world; ++r) acc += *(volatile double*)&me->slots[p][r][i];
        data[i] = bad ? __longlong_as_double(0x7ff8000000000000ll) : acc;
    }
    __syncthreads();
}

}  // namespace hw1f
