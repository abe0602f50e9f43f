// hw1f_multi.cu -- single-process multi-GPU front end of the C ABI (hw1f_multi_*): one engine per
// device, paths sharded by contiguous XORWOW subsequence range, ONE ncclAllReduce(ncclDouble, ncclSum)
// of the packed moment vector per workload, finalisation on device 0.  Used by the C++
// drivers (HW_GPUS=N); the Python/torchrun path uses one process per GPU and torch.distributed instead.
//
// NCCL is resolved with dlopen at hw1f_multi_create time, so libhw1f.so carries no NCCL dependency and
// never collides with the NCCL build a host application (e.g. PyTorch) has already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/hw1f.h"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err)
    {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !AllReduce || !GroupStart || !GroupEnd || !GetErrorString) {
            err = "libnccl lacks a required symbol";
            return false;
        }
        return true;
    }
};

}  // namespace

struct hw1f_multi {
    int n = 0;
    std::vector<hw1f_engine*> eng;
    std::vector<cudaStream_t> stream;
    std::vector<ncclComm_t> comm;
    std::vector<double*> d_mom;
    size_t mom_cap = 0;
    NcclApi nccl;
    hw1f_params p{};
    bool has_model = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // device-0 timing of the multi-GPU calls
    std::string err;
};

namespace {

#define M_REQUIRE(m, cond, msg) do { if (!(cond)) { (m)->err = (msg); return HW1F_ERR_INVALID; } } while (0)
#define M_CUDA(m, call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { (m)->err = std::string(#call) + ": " + cudaGetErrorString(_e); return HW1F_ERR_CUDA; } } while (0)
#define M_ENG(m, d, call) do { int _s = (call); if (_s != HW1F_OK) { (m)->err = std::string("device ") + std::to_string(d) + ": " + hw1f_last_error((m)->eng[d]); return _s; } } while (0)

void shard(uint64_t n_total, int rank, int world, uint64_t* first, uint64_t* count)
{
    const uint64_t base = n_total / world, extra = n_total % world;
    *first = rank * base + ((uint64_t)rank < extra ? rank : extra);
    *count = base + ((uint64_t)rank < extra ? 1 : 0);
}

int ensure_moments(hw1f_multi* m, size_t count)
{
    if (count <= m->mom_cap) return HW1F_OK;
    for (int d = 0; d < m->n; ++d) {
        M_CUDA(m, cudaSetDevice(d));
        if (m->d_mom[d]) cudaFree(m->d_mom[d]);
        m->d_mom[d] = nullptr;
        M_CUDA(m, cudaMalloc(&m->d_mom[d], count * sizeof(double)));
    }
    m->mom_cap = count;
    return HW1F_OK;
}

// the path's single exchange step: sum the per-GPU moment vectors over NVLink, in place
int allreduce(hw1f_multi* m, size_t count)
{
    if (m->n == 1) return HW1F_OK;
    ncclResult_t r = m->nccl.GroupStart();
    for (int d = 0; d < m->n && r == ncclSuccess; ++d)
        r = m->nccl.AllReduce(m->d_mom[d], m->d_mom[d], count, ncclDouble, ncclSum, m->comm[d], m->stream[d]);
    ncclResult_t r2 = m->nccl.GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) { m->err = std::string("ncclAllReduce: ") + m->nccl.GetErrorString(r); return HW1F_ERR_CUDA; }
    return HW1F_OK;
}

struct Rngs {
    std::vector<hw1f_rng*> h;
    ~Rngs() { for (auto* r : h) hw1f_rng_destroy(r); }
};

int make_rngs(hw1f_multi* m, uint64_t seed, uint64_t n_total, uint64_t offset, Rngs* out)
{
    M_REQUIRE(m, n_total >= (uint64_t)m->n, "fewer paths than GPUs");
    for (int d = 0; d < m->n; ++d) {
        uint64_t first, cnt;
        shard(n_total, d, m->n, &first, &cnt);
        hw1f_rng* r = nullptr;
        if (hw1f_rng_create(seed, first, cnt, &r) != HW1F_OK) { m->err = "hw1f_rng_create failed"; return HW1F_ERR_INVALID; }
        hw1f_rng_seek(r, offset);
        out->h.push_back(r);
    }
    return HW1F_OK;
}

}  // namespace

extern "C" {

const char* hw1f_multi_last_error(const hw1f_multi* m) { return m ? m->err.c_str() : "null handle"; }

int hw1f_multi_create(int n_gpus, hw1f_multi** out)
{
    if (!out) return HW1F_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { cudaGetLastError(); return HW1F_ERR_NO_DEVICE; }
    if (n_gpus <= 0 || n_gpus > count) n_gpus = count;
    hw1f_multi* m = new (std::nothrow) hw1f_multi();
    if (!m) return HW1F_ERR_INVALID;
    m->n = n_gpus;
    m->eng.assign(n_gpus, nullptr);
    m->stream.assign(n_gpus, nullptr);
    m->comm.assign(n_gpus, nullptr);
    m->d_mom.assign(n_gpus, nullptr);
    for (int d = 0; d < n_gpus; ++d) {
        int s = hw1f_engine_create(d, &m->eng[d]);
        if (s != HW1F_OK) { hw1f_multi_destroy(m); return s; }
        cudaSetDevice(d);
        if (cudaStreamCreateWithFlags(&m->stream[d], cudaStreamNonBlocking) != cudaSuccess) { hw1f_multi_destroy(m); return HW1F_ERR_CUDA; }
        hw1f_engine_set_stream(m->eng[d], m->stream[d]);
    }
    cudaSetDevice(0);
    if (cudaEventCreate(&m->ev0) != cudaSuccess || cudaEventCreate(&m->ev1) != cudaSuccess) { hw1f_multi_destroy(m); return HW1F_ERR_CUDA; }
    if (n_gpus > 1) {
        std::string err;
        if (!m->nccl.load(err)) { std::fprintf(stderr, "hw1f_multi_create: %s\n", err.c_str()); hw1f_multi_destroy(m); return HW1F_ERR_UNSUPPORTED; }
        std::vector<int> devs(n_gpus);
        for (int d = 0; d < n_gpus; ++d) devs[d] = d;
        if (m->nccl.CommInitAll(m->comm.data(), n_gpus, devs.data()) != ncclSuccess) { hw1f_multi_destroy(m); return HW1F_ERR_CUDA; }
        // first collective sets up the NVLink channels: do it here, not inside a timed workload
        if (ensure_moments(m, 1024) != HW1F_OK) { hw1f_multi_destroy(m); return HW1F_ERR_CUDA; }
        for (int d = 0; d < n_gpus; ++d) { cudaSetDevice(d); cudaMemsetAsync(m->d_mom[d], 0, 1024 * sizeof(double), m->stream[d]); }
        if (allreduce(m, 256) != HW1F_OK) { hw1f_multi_destroy(m); return HW1F_ERR_CUDA; }
        for (int d = 0; d < n_gpus; ++d) { cudaSetDevice(d); cudaStreamSynchronize(m->stream[d]); }
    }
    *out = m;
    return HW1F_OK;
}

int hw1f_multi_destroy(hw1f_multi* m)
{
    if (!m) return HW1F_OK;
    for (int d = 0; d < m->n; ++d) {
        cudaSetDevice(d);
        if (m->stream[d]) cudaStreamSynchronize(m->stream[d]);
        if (m->comm[d] && m->nccl.CommDestroy) m->nccl.CommDestroy(m->comm[d]);
        if (m->d_mom[d]) cudaFree(m->d_mom[d]);
        if (m->eng[d]) hw1f_engine_destroy(m->eng[d]);
        if (m->stream[d]) cudaStreamDestroy(m->stream[d]);
    }
    if (m->n > 0) cudaSetDevice(0);
    if (m->ev0) cudaEventDestroy(m->ev0);
    if (m->ev1) cudaEventDestroy(m->ev1);
    if (m->nccl.lib) dlclose(m->nccl.lib);
    delete m;
    return HW1F_OK;
}

int hw1f_multi_device_count(const hw1f_multi* m, int* n)
{
    if (!m || !n) return HW1F_ERR_INVALID;
    *n = m->n;
    return HW1F_OK;
}

int hw1f_multi_set_mode(hw1f_multi* m, int mode)
{
    if (!m) return HW1F_ERR_INVALID;
    for (int d = 0; d < m->n; ++d) M_ENG(m, d, hw1f_engine_set_mode(m->eng[d], mode));
    return HW1F_OK;
}

int hw1f_multi_set_model(hw1f_multi* m, const hw1f_params* p)
{
    if (!m || !p) return HW1F_ERR_INVALID;
    for (int d = 0; d < m->n; ++d) M_ENG(m, d, hw1f_set_model(m->eng[d], p));
    m->p = *p;
    m->has_model = true;
    return HW1F_OK;
}

int hw1f_multi_bond_curve(hw1f_multi* m, uint64_t seed, uint64_t n_paths_total, uint64_t normal_offset, float* P,
                          float* f, float* P_se, float* wall_ms)
{
    if (!m || !P || !f) return HW1F_ERR_INVALID;
    M_REQUIRE(m, m->has_model, "hw1f_multi_set_model() has not been called");
    const size_t count = 2 * (size_t)m->p.n_mat;
    int s = ensure_moments(m, 4 * (size_t)m->p.n_mat + 64);
    if (s != HW1F_OK) return s;
    Rngs rngs;
    s = make_rngs(m, seed, n_paths_total, normal_offset, &rngs);
    if (s != HW1F_OK) return s;
    for (int d = 0; d < m->n; ++d) M_ENG(m, d, hw1f_rng_prepare(m->eng[d], rngs.h[d]));
    cudaEvent_t e0 = m->ev0, e1 = m->ev1;
    M_CUDA(m, cudaSetDevice(0));
    M_CUDA(m, cudaEventRecord(e0, m->stream[0]));
    for (int d = 0; d < m->n; ++d) M_ENG(m, d, hw1f_bond_curve_moments(m->eng[d], rngs.h[d], m->d_mom[d]));
    s = allreduce(m, count);
    if (s != HW1F_OK) return s;
    M_CUDA(m, cudaSetDevice(0));
    M_CUDA(m, cudaEventRecord(e1, m->stream[0]));
    M_ENG(m, 0, hw1f_bond_curve_finish(m->eng[0], m->d_mom[0], n_paths_total, P, f, P_se));
    for (int d = 1; d < m->n; ++d) { M_CUDA(m, cudaSetDevice(d)); M_CUDA(m, cudaStreamSynchronize(m->stream[d])); }
    if (wall_ms) { M_CUDA(m, cudaSetDevice(0)); M_CUDA(m, cudaEventElapsedTime(wall_ms, e0, e1)); }
    return HW1F_OK;
}

int hw1f_multi_zbc_cv(hw1f_multi* m, uint64_t seed, uint64_t n_paths_total, uint64_t normal_offset, float S1, float S2,
                      float K, const float* P_mkt, const float* f_mkt, int32_t n_steps_S1, hw1f_zbc_result* out)
{
    if (!m || !P_mkt || !f_mkt || !out) return HW1F_ERR_INVALID;
    M_REQUIRE(m, m->has_model, "hw1f_multi_set_model() has not been called");
    int s = ensure_moments(m, 4 * (size_t)m->p.n_mat + 64);
    if (s != HW1F_OK) return s;
    int32_t n = n_steps_S1;
    if (n < 0) M_ENG(m, 0, hw1f_steps_to(m->eng[0], S1, &n));
    Rngs rngs;
    s = make_rngs(m, seed, n_paths_total, normal_offset, &rngs);
    if (s != HW1F_OK) return s;
    for (int d = 0; d < m->n; ++d)
        M_ENG(m, d, hw1f_zbc_cv_moments(m->eng[d], rngs.h[d], S1, S2, K, P_mkt, f_mkt, n, m->d_mom[d]));
    s = allreduce(m, 5);
    if (s != HW1F_OK) return s;
    out->n_steps_S1 = n;
    M_ENG(m, 0, hw1f_zbc_cv_finish(m->eng[0], m->d_mom[0], n_paths_total, P_mkt[m->p.n_mat - 1], out));
    out->n_steps_S1 = n;
    for (int d = 1; d < m->n; ++d) { M_CUDA(m, cudaSetDevice(d)); M_CUDA(m, cudaStreamSynchronize(m->stream[d])); }
    return HW1F_OK;
}

int hw1f_multi_fused(hw1f_multi* m, uint64_t seed, uint64_t n_paths_total, uint64_t normal_offset, float S1, float S2,
                     float K, const float* P_mkt, const float* f_mkt, float eps, int32_t n_steps_S1, float* P,
                     float* f, float* P_se, hw1f_zbc_result* zbc, hw1f_vega_result* vega, float* wall_ms)
{
    if (!m || !P_mkt || !f_mkt || !P || !f || !zbc || !vega) return HW1F_ERR_INVALID;
    M_REQUIRE(m, m->has_model, "hw1f_multi_set_model() has not been called");
    M_REQUIRE(m, eps > 0.0f, "hw1f_multi_fused needs eps > 0");
    const size_t count = 2 * (size_t)m->p.n_mat + HW1F_FUSED_EXTRA + HW1F_FUSED_FD_EXTRA;
    int s = ensure_moments(m, 4 * (size_t)m->p.n_mat + 64);
    if (s != HW1F_OK) return s;
    int32_t n = n_steps_S1;
    if (n < 0) M_ENG(m, 0, hw1f_steps_to(m->eng[0], S1, &n));
    Rngs rngs;
    s = make_rngs(m, seed, n_paths_total, normal_offset, &rngs);
    if (s != HW1F_OK) return s;
    for (int d = 0; d < m->n; ++d) M_ENG(m, d, hw1f_rng_prepare(m->eng[d], rngs.h[d]));
    cudaEvent_t e0 = m->ev0, e1 = m->ev1;
    M_CUDA(m, cudaSetDevice(0));
    M_CUDA(m, cudaEventRecord(e0, m->stream[0]));
    for (int d = 0; d < m->n; ++d)
        M_ENG(m, d, hw1f_fused_fd_moments(m->eng[d], rngs.h[d], S1, S2, K, P_mkt, f_mkt, eps, n, m->d_mom[d]));
    s = allreduce(m, count);
    if (s != HW1F_OK) return s;
    M_CUDA(m, cudaSetDevice(0));
    M_CUDA(m, cudaEventRecord(e1, m->stream[0]));
    M_ENG(m, 0, hw1f_fused_finish(m->eng[0], m->d_mom[0], n_paths_total, P_mkt[m->p.n_mat - 1], eps, n, P, f, P_se, zbc,
                                  vega));
    for (int d = 1; d < m->n; ++d) { M_CUDA(m, cudaSetDevice(d)); M_CUDA(m, cudaStreamSynchronize(m->stream[d])); }
    float ms = 0.f;
    M_CUDA(m, cudaSetDevice(0));
    M_CUDA(m, cudaEventElapsedTime(&ms, e0, e1));
    vega->ms_pathwise = vega->ms_fd = ms;
    if (wall_ms) *wall_ms = ms;
    return HW1F_OK;
}

int hw1f_multi_vega_pathwise(hw1f_multi* m, uint64_t seed, uint64_t n_paths_total, uint64_t normal_offset, float S1,
                             float S2, float K, const float* P_mkt, const float* f_mkt, int32_t n_steps_S1,
                             double* vega, double* vega_se)
{
    if (!m || !P_mkt || !f_mkt || !vega) return HW1F_ERR_INVALID;
    M_REQUIRE(m, m->has_model, "hw1f_multi_set_model() has not been called");
    int s = ensure_moments(m, 4 * (size_t)m->p.n_mat + 64);
    if (s != HW1F_OK) return s;
    int32_t n = n_steps_S1;
    if (n < 0) M_ENG(m, 0, hw1f_steps_to(m->eng[0], S1, &n));
    Rngs rngs;
    s = make_rngs(m, seed, n_paths_total, normal_offset, &rngs);
    if (s != HW1F_OK) return s;
    for (int d = 0; d < m->n; ++d)
        M_ENG(m, d, hw1f_vega_pathwise_moments(m->eng[d], rngs.h[d], S1, S2, K, P_mkt, f_mkt, n, m->d_mom[d]));
    s = allreduce(m, 2);
    if (s != HW1F_OK) return s;
    double mom[2];
    M_CUDA(m, cudaSetDevice(0));
    M_CUDA(m, cudaMemcpyAsync(mom, m->d_mom[0], sizeof(mom), cudaMemcpyDeviceToHost, m->stream[0]));
    for (int d = 0; d < m->n; ++d) { M_CUDA(m, cudaSetDevice(d)); M_CUDA(m, cudaStreamSynchronize(m->stream[d])); }
    const double np = (double)n_paths_total;
    *vega = mom[0] / np;
    if (vega_se) {
        const double var = (np > 1) ? (mom[1] - mom[0] * mom[0] / np) / (np - 1.0) : 0.0;
        *vega_se = (var > 0) ? sqrt(var / np) : 0.0;
    }
    return HW1F_OK;
}

}  // extern "C"
