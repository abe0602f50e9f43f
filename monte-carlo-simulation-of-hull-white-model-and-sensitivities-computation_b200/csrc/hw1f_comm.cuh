// hw1f_comm.cuh -- device side of the path's single exchange step: an all-reduce of the moment vector
// over NVLink peer memory (CUDA IPC mailboxes), shared by the stand-alone kernel (hw1f_comm.cu) and by
// the tail kernel behind every simulation launch (hw1f_tail.cuh), where the block that closed the
// reduction posts the vector to the peers itself -- no separate launch between reduction and collective.
#pragma once
#include <cuda_runtime.h>

namespace hw1f {

constexpr int kCommMaxWorld = 8;
constexpr int kCommMaxCount = 512;             // doubles per collective (Q1: 202, fused: 220, recalibration curves: 404)
constexpr unsigned kCommSpinLimit = 20000000u;  // a few seconds of polling: far beyond any healthy skew, still finite

// Mailbox: per epoch parity and source rank, TWO 64-bit words per double -- (epoch << 32 | low half) and
// (epoch << 32 | high half).  Every 8-byte word carries its own validity flag and 8-byte stores are delivered whole, so
// a receiver can consume a value the moment both words show the current epoch: no fence, no separate flag, ONE NVLink
// one-way trip per exchange (the idea of NCCL's low-latency "LL" protocol; round 1 and the first round-2 form posted
// plain doubles, fenced at system scope and then raised a flag: two more trips on the critical path).
struct Mailbox {
    unsigned long long ll[2][kCommMaxWorld][2 * kCommMaxCount];
    unsigned timeouts;
};

struct CommDev {
    Mailbox* peer[kCommMaxWorld];   // peer[r] = rank r's mailbox as mapped into this process
    int rank, world;                // world <= 1: no exchange
};

__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// In-place SUM all-reduce of data[0, count) by ONE block (every thread of the block calls this; count <=
// kCommMaxCount; data[] complete and visible to the block; epoch != 0):
//   1. every thread posts its values, two flagged words each, into this rank's slot of EVERY rank's mailbox (8-byte
//      stores over NVLink / NVSwitch, nothing else);
//   2. every thread polls the LOCAL mailbox for its values from every rank (bounded spin) and
//   3. sums them in RANK ORDER: bit-identical on every rank and from run to run.
// Mailboxes are double-buffered by epoch parity: a rank reaches epoch e+2 only after every peer has posted e+1, i.e.
// after every peer finished reading e.  A lost peer poisons the WHOLE vector with NaN (every *_finish rejects it) and
// bumps the mailbox's time-out counter instead of hanging the GPU.
__device__ __forceinline__ void block_peer_allreduce(const CommDev& c, double* __restrict__ data, int count, unsigned epoch)
{
    __shared__ int s_timed_out;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int p = epoch & 1u;
    const unsigned long long tag = (unsigned long long)epoch << 32;
    if (tid == 0) s_timed_out = 0;
    __syncthreads();
    for (int i = tid; i < count; i += nthr) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(data[i]);
        const unsigned long long lo = tag | (bits & 0xffffffffull), hi = tag | (bits >> 32);
        for (int r = 0; r < c.world; ++r) {
            unsigned long long* slot = &c.peer[r]->ll[p][c.rank][2 * i];
            st_relaxed_sys_u64(slot, lo);
            st_relaxed_sys_u64(slot + 1, hi);
        }
    }
    Mailbox* me = c.peer[c.rank];
    for (int i = tid; i < count; i += nthr) {
        double acc = 0.0;
        bool bad = false;
        for (int r = 0; r < c.world; ++r) {
            const unsigned long long* slot = &me->ll[p][r][2 * i];
            unsigned long long lo = ld_relaxed_sys_u64(slot), hi = ld_relaxed_sys_u64(slot + 1);
            unsigned spins = 0;
            while ((unsigned)(lo >> 32) != epoch || (unsigned)(hi >> 32) != epoch) {
                if (++spins > kCommSpinLimit) { bad = true; break; }
                __nanosleep(32);
                lo = ld_relaxed_sys_u64(slot);
                hi = ld_relaxed_sys_u64(slot + 1);
            }
            if (bad) break;
            acc += __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
        }
        if (bad) s_timed_out = 1;
        data[i] = acc;
    }
    __syncthreads();
    if (s_timed_out) {   // one lost value poisons the whole vector: no partially reduced results
        if (tid == 0) atomicAdd(&me->timeouts, 1u);
        for (int i = tid; i < count; i += nthr) data[i] = __longlong_as_double(0x7ff8000000000000ll);
        __syncthreads();
    }
}

}  // namespace hw1f
