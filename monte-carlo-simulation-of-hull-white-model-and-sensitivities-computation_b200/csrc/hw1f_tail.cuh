// hw1f_tail.cuh -- everything that follows a simulation launch, as ONE small kernel: the second and third level of
// the deterministic reduction tree, the moment all-reduce with the peers, and the path-independent epilogues.
//
// Round 1 ran these as separate launches after every simulation (reduce_curve_kernel / reduce_partials_kernel,
// peer_allreduce_kernel, curve_epilogue_kernel, bond_plan_kernel, a device->host copy + two synchronisations):
// 20-50 us of launch gaps and round trips on a 0.3-0.6 ms pass.  tail_kernel is launched with programmatic stream
// serialisation right behind the simulation kernel (its blocks are resident and parked in griddepcontrol.wait when the
// last simulation block retires):
//   level 1  (simulation kernel) every block writes partials[run][block][nq] in double;
//   level 2  tail block (g, run) sums the rows of group g -- kTailGroup consecutive simulation blocks -- in block
//            order into gpart[run][g][nq], fences, takes a ticket of the run;
//   level 3  the block that takes the run's last ticket sums the group rows in group order, undoes the centring of
//            the curve sums (double) and stores moments[run][..] -- fixed summation order at every level, independent
//            of which block arrives when: bit-reproducible, no floating-point atomics;
//   then     the same block (optionally) all-reduces the vector with the peers over NVLink (hw1f_comm.cuh: it posts
//            the vector to the peers' mailboxes itself, no launch between reduction and exchange), runs the curve
//            epilogue (P, f, standard error) and the bond plans of recalibrated curves, and stores moments and curves
//            straight into mapped pinned host memory: the host only waits for the stream.
// Ticket counters reset themselves, so consecutive launches need no memset.
//
// (The tail was first built INTO the simulation kernels -- last block done -- and measured: ptxas then re-schedules
// the simulation loop of the calling kernel (same 258 instructions, different order, also when the tail is a
// __noinline__ call) and the Q1 kernel loses 8 %, 571 -> 622 us; profiles/r02_ab_tail_in_kernel.txt.)
#pragma once
#include "hw1f_comm.cuh"
#include "hw1f_kernels.cuh"

namespace hw1f {

constexpr int kTailGroup = 32;

struct TailArgs {
    unsigned* counters;            // [n_runs], zero outside launches
    double* gpart;                 // [n_runs][n_groups][nq]
    double* moments;               // [n_runs][out_stride] device result
    double* host_mom;              // same layout in mapped pinned host memory, or null
    int out_stride;
    int n_ext_out;                 // number of non-curve moments emitted (<= nq - nqc)
    int ncur;                      // centred curve scenarios at the front of the vector (0, 1, 2)
    int n_mat;
    const float* center0;          // centring constants of scenario 0 / 1: c_m = center_scale * center[m]
    const float* center1;
    float center_scale;
    unsigned long long n_local;    // subsequences of this launch (un-centring)
    // exchange (single run only): world <= 1 = none
    CommDev comm;
    unsigned epoch;
    // curve epilogue after the (all-)reduction: P, f (and P_se) of every curve scenario
    int epi;                       // 0 none, 1 run it
    unsigned long long n_total;    // subsequences over all ranks
    float inv_dT;
    float* dev_curve;              // [ncur][2][n_mat] P, f on the device (market curves of a later pricing), or null
    float* host_curve;             // [ncur][3][n_mat] P, f, P_se in mapped pinned host memory, or null
    PlanJob plan;                  // bond plans on dev_curve (recalibrated FD), n_scen = 0: none
};

// What follows the reduction, by ONE block per run: exchange with the peers, copy of the moment vector to the host,
// curve epilogue, bond plans on the fresh curves.  moments[run][..] must be complete and visible to the block.
__device__ __forceinline__ void tail_publish(const TailArgs& ta, const ModelDev& md, int run, int nqc, float* scratch)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int nm = ta.n_mat;
    const int n_out = nqc + ta.n_ext_out;
    double* mom = ta.moments + (size_t)run * ta.out_stride;

    // ---- exchange: this block posts the vector to the peers itself ----
    if (ta.comm.world > 1 && run == 0) block_peer_allreduce(ta.comm, mom, n_out, ta.epoch);

    if (ta.host_mom) {
        double* hm = ta.host_mom + (size_t)run * ta.out_stride;
        for (int q = tid; q < n_out; q += nthr) hm[q] = mom[q];
    }
    if (ta.epi) {
        for (int s = 0; s < ta.ncur; ++s) {
            float* hP = ta.host_curve ? ta.host_curve + ((size_t)run * ta.ncur + s) * 3 * nm : nullptr;
            float* dP = ta.dev_curve ? ta.dev_curve + ((size_t)run * ta.ncur + s) * 2 * nm : nullptr;
            // P and f go to the device curve if one is asked for (market curves of a later pricing), else to the host
            float* outP = dP ? dP : hP;
            float* outf = dP ? dP + nm : hP + nm;
            curve_epilogue_block(mom + (size_t)s * 2 * nm, nm, ta.n_total, ta.inv_dT, outP, outf, hP ? hP + 2 * nm : nullptr,
                                 scratch);
            if (dP && hP)
                for (int m = tid; m < nm; m += nthr) { hP[m] = dP[m]; hP[nm + m] = dP[nm + m]; }
        }
        if (ta.plan.n_scen > 0) {
            __syncthreads();
            run_plan_job(md, ta.plan, tid);
        }
    }
}

// Column sums of a row-major [n_rows][nq] double matrix in global memory by one block, in a FIXED order (so the result
// does not depend on anything but the data): the rows are split into `parts` contiguous ranges summed by different
// threads -- up to eight independent loads in flight per thread, latency of a few round trips instead of n_rows --
// and the range sums are added in range order.  tot[q] (shared, nq doubles) receives the sums; dscr: parts * nq
// doubles of shared scratch.  All threads of the block call it.
__device__ __forceinline__ void column_sums(const double* __restrict__ rows, int n_rows, int nq, double* tot, double* dscr,
                                            int max_parts)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    int parts = nthr / nq;
    if (parts > max_parts) parts = max_parts;
    if (parts > n_rows) parts = n_rows;
    if (parts < 1) parts = 1;
    const int per = (n_rows + parts - 1) / parts;
    auto range_sum = [&](int q, int r0, int r1) {
        double acc = 0.0;
        for (int r = r0; r < r1; r += 8) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (r + j < r1) ? __ldcg(rows + (size_t)(r + j) * nq + q) : 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j];
        }
        return acc;
    };
    if (parts == 1) {
        for (int q = tid; q < nq; q += nthr) tot[q] = range_sum(q, 0, n_rows);
        __syncthreads();
        return;
    }
    const int part = tid / nq, q = tid - part * nq;
    if (part < parts) {
        const int r0 = part * per, r1 = min(r0 + per, n_rows);
        dscr[part * nq + q] = range_sum(q, r0, r1);
    }
    __syncthreads();
    if (tid < nq) {
        double acc = dscr[tid];
        for (int p = 1; p < parts; ++p) acc += dscr[p * nq + tid];
        tot[tid] = acc;
    }
    __syncthreads();
}

constexpr int kTailThreads = 1024;

// grid (n_groups, n_runs), kTailThreads threads, dynamic shared memory: tail_smem_bytes(nq, n_mat)
__global__ void __launch_bounds__(kTailThreads)
tail_kernel(TailArgs ta, ModelDev md, const double* __restrict__ partials, int n_blocks, int nq, int nqc)
{
    extern __shared__ __align__(16) double tail_scratch[];
    __shared__ int s_last;
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the simulation launch in front has completed and flushed
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n_groups = gridDim.x, g = blockIdx.x, run = blockIdx.y;
    const int g_first = g * kTailGroup;
    const int g_count = min(kTailGroup, n_blocks - g_first);
    double* tot = tail_scratch;          // [nq]
    double* dscr = tail_scratch + nq;    // [parts][nq]
    const int max_parts = 8;

    // ---- level 2: the rows of this group ----
    column_sums(partials + ((size_t)run * n_blocks + g_first) * nq, g_count, nq, tot, dscr, max_parts);
    {
        double* gp = ta.gpart + ((size_t)run * n_groups + g) * nq;
        for (int q = tid; q < nq; q += nthr) gp[q] = tot[q];
    }
    __threadfence();   // the group row is visible device-wide before the ticket is
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&ta.counters[run], 1u) == (unsigned)(n_groups - 1));
    __syncthreads();
    if (!s_last) return;

    // ---- level 3: this block closes the run ----
    __threadfence();
    if (tid == 0) ta.counters[run] = 0;
    column_sums(ta.gpart + (size_t)run * n_groups * nq, n_groups, nq, tot, dscr, max_parts);
    double* mom = ta.moments + (size_t)run * ta.out_stride;
    const int nm = ta.n_mat;
    for (int idx = tid; idx < ta.ncur * nm; idx += nthr) {
        // sum p0 = sum d + n c,  sum p0^2 = sum d^2 + 2 c sum d + n c^2   (d = p0 - c, c the noise-free value)
        const int s = idx / nm, m = idx - s * nm;
        const int q0 = s * 2 * nm + m;
        double o0 = 0.0, o1 = 0.0;
        if (m > 0) {
            const double sd = tot[q0], sdd = tot[q0 + nm];
            const double c = (double)ta.center_scale * (double)(s ? ta.center1 : ta.center0)[m], n = (double)ta.n_local;
            o0 = sd + n * c;
            o1 = sdd + 2.0 * c * sd + n * c * c;
        }
        mom[q0] = o0;
        mom[q0 + nm] = o1;
    }
    for (int k = tid; k < ta.n_ext_out; k += nthr) mom[nqc + k] = tot[nqc + k];
    __syncthreads();
    tail_publish(ta, md, run, nqc, reinterpret_cast<float*>(tail_scratch));
}

// tot[nq] + up to 8 range sums [8][nq] when nq is small against the block, and n_mat floats for the epilogue
inline size_t tail_smem_bytes(int nq, int n_mat)
{
    int parts = kTailThreads / nq;
    if (parts > 8) parts = 8;
    if (parts < 1) parts = 1;
    const size_t a = (size_t)(1 + (parts > 1 ? parts : 0)) * nq * sizeof(double), b = (size_t)n_mat * sizeof(float);
    return (a > b ? a : b) + 16;
}

// ---- confidence intervals of f(0,T) and theta(T): batch means over the simulation blocks --------------------------
// The curve kernels keep sum d and sum d^2 per maturity (d = p0 - c_m), which gives the standard error of P but not of
// quantities that DIFFERENCE neighbouring maturities (f = -d ln P / dT, theta = df/dT + a f + ...): those need the
// covariance of P at nearby maturities.  Every simulation block is an independent batch of subsequences, so the
// covariance band comes from the per-block partial sums the reduction tree already holds, at no cost to the hot loop:
//   C[l][m] = sum_b (S_b[m] - n_b mu_m)(S_b[m+l] - n_b mu_(m+l)),  l = 0..kCiLags-1,
// S_b[m] = partials[b][m] (block sum of d_m), n_b = subsequences block b simulated, mu_m = sum_b S_b[m] / n.
// Also stores tot[m] = sum_b S_b[m].  Grid: n_mat blocks of 256 threads.  out: [kCiLags + 1][n_mat] doubles.
constexpr int kCiLags = 5;

__device__ __forceinline__ unsigned long long block_subsequences(const StreamGeom& g, unsigned b, unsigned grid)
{
    unsigned long long n = 0;
    for (unsigned long long c = b; c < g.n_chunks; c += grid) {
        const unsigned long long lo = (g.chunk0 + c) << kChunkLog2, hi = lo + kChunk;
        const unsigned long long a = lo > g.first_path ? lo : g.first_path;
        const unsigned long long e = hi < g.first_path + g.n_paths ? hi : g.first_path + g.n_paths;
        if (e > a) n += e - a;
    }
    return n;
}

__global__ void __launch_bounds__(256)
curve_batch_cov_kernel(const double* __restrict__ partials, int n_blocks, int stride, int n_mat, StreamGeom g,
                       double* __restrict__ out)
{
    __shared__ double sh[kCiLags + 1][256];
    const int m = blockIdx.x, tid = threadIdx.x;
    const double n_all = (double)g.n_paths;
    // pass 1: totals of columns m .. m+kCiLags-1
    double t[kCiLags];
#pragma unroll
    for (int l = 0; l < kCiLags; ++l) t[l] = 0.0;
    for (int b = tid; b < n_blocks; b += 256)
#pragma unroll
        for (int l = 0; l < kCiLags; ++l)
            if (m + l < n_mat) t[l] += partials[(size_t)b * stride + m + l];
#pragma unroll
    for (int l = 0; l < kCiLags; ++l) sh[l][tid] = t[l];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o)
#pragma unroll
            for (int l = 0; l < kCiLags; ++l) sh[l][tid] += sh[l][tid + o];
        __syncthreads();
    }
    double mu[kCiLags];
#pragma unroll
    for (int l = 0; l < kCiLags; ++l) mu[l] = sh[l][0] / n_all;
    const double tot_m = sh[0][0];
    __syncthreads();
    // pass 2: centred cross products
    double c[kCiLags];
#pragma unroll
    for (int l = 0; l < kCiLags; ++l) c[l] = 0.0;
    for (int b = tid; b < n_blocks; b += 256) {
        const double nb = (double)block_subsequences(g, (unsigned)b, (unsigned)n_blocks);
        const double d0 = partials[(size_t)b * stride + m] - nb * mu[0];
#pragma unroll
        for (int l = 0; l < kCiLags; ++l)
            if (m + l < n_mat) c[l] += d0 * (partials[(size_t)b * stride + m + l] - nb * mu[l]);
    }
#pragma unroll
    for (int l = 0; l < kCiLags; ++l) sh[l][tid] = c[l];
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o)
#pragma unroll
            for (int l = 0; l < kCiLags; ++l) sh[l][tid] += sh[l][tid + o];
        __syncthreads();
    }
    if (tid < kCiLags) out[(size_t)tid * n_mat + m] = sh[tid][0];
    if (tid == 0) out[(size_t)kCiLags * n_mat + m] = tot_m;
}

// the same publication step as a launch of its own: for moment vectors that were reduced by separate kernels
// (reference-order mode) or that come back from an external all-reduce (hw1f_*_finish).  One block per run.
__global__ void __launch_bounds__(256) tail_publish_kernel(TailArgs ta, ModelDev md, int nqc)
{
    extern __shared__ float publish_scratch[];   // n_mat floats
    tail_publish(ta, md, blockIdx.x, nqc, publish_scratch);
}

}  // namespace hw1f
