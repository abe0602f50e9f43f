// hw1f_comm.cu -- host side of the path's single exchange step (device side: hw1f_comm.cuh).
//
// One process per GPU (torchrun).  Every rank owns a small mailbox in device memory that all peers map through CUDA
// IPC.  The all-reduce of the moment vector (<= 512 doubles) is done by ONE block: it posts the vector into its slot of
// EVERY peer's mailbox as flagged 8-byte words (plain stores over NVLink / NVSwitch; every word carries the epoch, so
// there is no fence and no separate flag), polls its own mailbox until every rank's words of this epoch have arrived
// (bounded spin: a lost peer poisons the result with NaN and bumps a counter instead of hanging the GPU), and sums the
// slots in RANK ORDER, so the result is bit-identical on every rank and from run to run (NCCL's ring/tree order
// depends on the communicator).  Mailboxes are double-buffered by epoch parity.
//
// hw1f_comm_attach hands the mailboxes to the engine: the tail kernel behind every *_moments simulation launch then
// does the exchange itself (no launch between reduction and collective).  hw1f_comm_allreduce is the same device code
// as a launch of its own.  For a 1.6 KB payload the collective is pure latency: one NVLink one-way trip here, ~23 us
// for NCCL's all-reduce.  NCCL (torch.distributed) remains the plumbing for the handle exchange and the verification
// path in bench.py.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/hw1f.h"
#include "hw1f_comm.cuh"

using namespace hw1f;

// the engine side of hw1f_comm_attach (hw1f_api.cu): hands the mapped mailboxes and the epoch counter to the
// engine, whose *_moments entry points then all-reduce in the tail of their simulation kernel
extern "C" int hw1f_engine_attach_comm_internal(hw1f_engine* e, const void* comm_dev, unsigned* epoch, int on);

namespace {

// the exchange as a launch of its own (hw1f_comm_allreduce): one block, device code shared with the kernel tails
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(CommDev c, double* __restrict__ data, int count, unsigned epoch)
{
    block_peer_allreduce(c, data, count, epoch);
}

}  // namespace

struct hw1f_comm {
    hw1f_engine* eng = nullptr;
    int device = 0, rank = -1, world = 0;
    Mailbox* local = nullptr;
    CommDev dev{};
    bool opened[kCommMaxWorld] = {false};
    bool attached = false;
    unsigned epoch = 0;
    cudaStream_t stream = nullptr;
    std::string err;
};

extern "C" {

const char* hw1f_comm_last_error(const hw1f_comm* c) { return c ? c->err.c_str() : "null comm"; }

int hw1f_comm_create(hw1f_engine* eng, int world, void* ipc_handle64, hw1f_comm** out)
{
    if (!eng || !ipc_handle64 || !out || world < 1 || world > kCommMaxWorld) return HW1F_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    hw1f_comm* c = new (std::nothrow) hw1f_comm();
    if (!c) return HW1F_ERR_INVALID;
    c->eng = eng;
    c->world = world;
    hw1f_engine_device(eng, &c->device);
    cudaSetDevice(c->device);
    cudaError_t e = cudaMalloc(&c->local, sizeof(Mailbox));
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, sizeof(Mailbox));
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) {
        std::fprintf(stderr, "hw1f_comm_create: %s\n", cudaGetErrorString(e));
        if (c->local) cudaFree(c->local);
        delete c;
        return HW1F_ERR_CUDA;
    }
    memcpy(ipc_handle64, &h, 64);
    *out = c;
    return HW1F_OK;
}

int hw1f_comm_connect(hw1f_comm* c, int rank, const void* all_handles, void* cuda_stream)
{
    if (!c || !all_handles || rank < 0 || rank >= c->world) return HW1F_ERR_INVALID;
    cudaSetDevice(c->device);
    c->rank = rank;
    c->stream = (cudaStream_t)cuda_stream;
    c->dev.rank = rank;
    c->dev.world = c->world;
    for (int r = 0; r < c->world; ++r) {
        if (r == rank) { c->dev.peer[r] = c->local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            c->err = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e);
            cudaGetLastError();
            return HW1F_ERR_CUDA;
        }
        c->dev.peer[r] = (Mailbox*)p;
        c->opened[r] = true;
    }
    return HW1F_OK;
}

int hw1f_comm_allreduce(hw1f_comm* c, double* d_data, int32_t count)
{
    if (!c || !d_data || c->rank < 0) return HW1F_ERR_INVALID;
    if (count < 1 || count > kCommMaxCount) { c->err = "count must be in [1,512]"; return HW1F_ERR_INVALID; }
    cudaSetDevice(c->device);
    if (c->epoch == 0xffffffffu) { c->err = "communicator exhausted (2^32 exchanges): create a new one"; return HW1F_ERR_COMM; }
    ++c->epoch;   // shared with the engine's tail kernels when attached: one sequence of epochs per communicator
    peer_allreduce_kernel<<<1, 256, 0, c->stream>>>(c->dev, d_data, count, c->epoch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { c->err = std::string("peer_allreduce_kernel: ") + cudaGetErrorString(e); return HW1F_ERR_CUDA; }
    return HW1F_OK;
}

int hw1f_comm_attach(hw1f_comm* c, int on)
{
    if (!c || c->rank < 0) return HW1F_ERR_INVALID;
    const int st = hw1f_engine_attach_comm_internal(c->eng, &c->dev, &c->epoch, on);
    if (st == HW1F_OK) c->attached = on != 0;
    return st;
}

int hw1f_comm_timeouts(hw1f_comm* c, uint32_t* n)
{
    if (!c || !n) return HW1F_ERR_INVALID;
    cudaSetDevice(c->device);
    cudaError_t e = cudaMemcpy(n, &c->local->timeouts, sizeof(unsigned), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return HW1F_ERR_CUDA; }
    return HW1F_OK;
}

int hw1f_comm_destroy(hw1f_comm* c)
{
    if (!c) return HW1F_OK;
    cudaSetDevice(c->device);
    if (c->attached) hw1f_engine_attach_comm_internal(c->eng, &c->dev, &c->epoch, 0);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->dev.peer[r]);
    if (c->local) cudaFree(c->local);
    delete c;
    return HW1F_OK;
}

}  // extern "C"
