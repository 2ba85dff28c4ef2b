// Centroid update (balancekmeans/__init__.py:314-324, :343-346) as a DETERMINISTIC segmented
// reduction: no floating-point atomics anywhere, so the same assignment always gives the same
// centroids bit for bit, on any grid.
//
//   1. member lists: a stable counting sort of job indices by cluster (integer atomics only;
//      ranks come from warp match/ballot + a fixed prefix over blocks), so each cluster's members
//      are listed in ascending job order.
//   2. partial sums: CTA (cluster c, split s) adds the rows of members s, s+SPLITS, ... in that
//      order; one thread owns 4 consecutive dims (float4, coalesced 2 KB row reads), 8 rows in
//      flight per thread.  X is read exactly once: 4*D bytes per vector, HBM-bound.
//   3. combine: the SPLITS partials are added in index order; mean = sum / count; the centre
//      shift sum_k |c_k - c_k_prev|_2 is reduced in a fixed order by one CTA.
// Multi-GPU: step 2's [K][D] sums and [K] counts are what ranks allreduce (NCCL) before step 3.
#include "common.cuh"

namespace rqk {

constexpr int CU_BLOCK_ROWS = 1024;   // rows per CTA in the sort kernels (32 warps x 32 rows)
constexpr int CU_SPLITS = 16;

// blockcnt[b][k] = number of rows of block b assigned to cluster k
__global__ void __launch_bounds__(1024)
cu_block_count_kernel(const int* __restrict__ assign, long long n, int K, int* __restrict__ blockcnt) {
    extern __shared__ int scnt[];   // [K]
    for (int i = threadIdx.x; i < K; i += blockDim.x) scnt[i] = 0;
    __syncthreads();
    long long row = (long long)blockIdx.x * CU_BLOCK_ROWS + threadIdx.x;
    if (row < n) atomicAdd(&scnt[assign[row]], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) blockcnt[(long long)blockIdx.x * K + i] = scnt[i];
}

// per cluster: exclusive prefix over blocks (in place) and total count.  Grid = K CTAs x 256 threads.
__global__ void __launch_bounds__(256)
cu_block_scan_kernel(int* __restrict__ blockcnt, int nblocks, int K, int* __restrict__ counts) {
    const int k = blockIdx.x, tid = threadIdx.x;
    __shared__ int part[256];
    // each thread owns a contiguous run of blocks
    int per = (nblocks + 255) / 256;
    int b0 = tid * per, b1 = min(b0 + per, nblocks);
    int s = 0;
    for (int b = b0; b < b1; ++b) s += blockcnt[(long long)b * K + k];
    part[tid] = s;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        int v = (tid >= d) ? part[tid - d] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int run = part[tid] - s;
    for (int b = b0; b < b1; ++b) {
        int c = blockcnt[(long long)b * K + k];
        blockcnt[(long long)b * K + k] = run;
        run += c;
    }
    if (tid == 255) counts[k] = part[255];
}

// offsets[k] = exclusive prefix of counts (single CTA, K <= 1024)
__global__ void cu_offsets_kernel(const int* __restrict__ counts, int K, long long* __restrict__ offsets) {
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int k = 0; k < K; ++k) { offsets[k] = run; run += counts[k]; }
        offsets[K] = run;
    }
}

// order[offsets[k] + rank] = row, rank = #rows before `row` (ascending) with the same cluster
__global__ void __launch_bounds__(1024)
cu_scatter_kernel(const int* __restrict__ assign, long long n, int K, const int* __restrict__ blockpre,
                  const long long* __restrict__ offsets, int* __restrict__ order) {
    extern __shared__ int wcnt[];   // [32 warps][K]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * K; i += blockDim.x) wcnt[i] = 0;
    __syncthreads();
    long long row = (long long)blockIdx.x * CU_BLOCK_ROWS + tid;
    int c = (row < n) ? assign[row] : -1;
    unsigned int peers = __match_any_sync(0xffffffffu, c);
    int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (c >= 0 && rank_in_warp == 0) wcnt[warp * K + c] = __popc(peers);
    __syncthreads();
    if (c >= 0) {
        int before = 0;
        for (int w = 0; w < warp; ++w) before += wcnt[w * K + c];
        long long pos = offsets[c] + blockpre[(long long)blockIdx.x * K + c] + before + rank_in_warp;
        order[pos] = (int)row;
    }
}

// partial[k][s][dim]: sum of the member rows s, s+SPLITS, ... of cluster k.  blockDim = dim/4.
__global__ void cu_partial_sum_kernel(const float* __restrict__ x, int dim, const int* __restrict__ order,
                                      const long long* __restrict__ offsets, float* __restrict__ partial) {
    const int k = blockIdx.x, s = blockIdx.y, t = threadIdx.x;
    const long long beg = offsets[k], end = offsets[k + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    long long m = beg + s;
    constexpr int U = 8;
    for (; m + (long long)(U - 1) * CU_SPLITS < end; m += (long long)U * CU_SPLITS) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int row = order[m + (long long)u * CU_SPLITS];
            v[u] = *reinterpret_cast<const float4*>(x + (long long)row * dim + t * 4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; m < end; m += CU_SPLITS) {
        int row = order[m];
        float4 v = *reinterpret_cast<const float4*>(x + (long long)row * dim + t * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(partial + ((long long)k * CU_SPLITS + s) * dim + t * 4) = acc;
}

// sums[k][dim] = partial[k][0] + partial[k][1] + ... in index order
__global__ void cu_combine_kernel(const float* __restrict__ partial, int dim, float* __restrict__ sums) {
    const int k = blockIdx.x;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        float a = 0.f;
#pragma unroll
        for (int s = 0; s < CU_SPLITS; ++s) a += partial[((long long)k * CU_SPLITS + s) * dim + d];
        sums[(long long)k * dim + d] = a;
    }
}

// centers[k] = sums[k] / counts[k] (empty clusters keep their old centre and are flagged), and
// shift_out[0] = sum_k |new_k - old_k|_2, shift_out[1] = number of empty clusters (as float).
// One CTA of 1024 threads, warp per cluster, fixed reduction order.
__global__ void __launch_bounds__(1024)
cu_finalize_kernel(const float* __restrict__ sums, const long long* __restrict__ counts, int K, int dim,
                   float* __restrict__ centers, float* __restrict__ shift_out, int* __restrict__ empty_mask) {
    __shared__ float snorm[1024];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 1024; i += 1024) snorm[i] = 0.f;
    __syncthreads();
    for (int k = warp; k < K; k += 32) {
        long long cnt = counts[k];
        float acc = 0.f;
        if (cnt > 0) {
            float inv_n = (float)cnt;
            for (int d = lane; d < dim; d += 32) {
                float nv = sums[(long long)k * dim + d] / inv_n;
                float dv = nv - centers[(long long)k * dim + d];
                centers[(long long)k * dim + d] = nv;
                acc = fmaf(dv, dv, acc);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            snorm[k] = sqrtf(acc);
            if (empty_mask) empty_mask[k] = (cnt == 0) ? 1 : 0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        int ne = 0;
        for (int k = 0; k < K; ++k) { s += snorm[k]; ne += (counts[k] == 0); }
        shift_out[0] = s;
        shift_out[1] = (float)ne;
    }
}

__global__ void cu_counts_to_i64_kernel(const int* __restrict__ c32, int K, long long* __restrict__ c64) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) c64[i] = c32[i];
}

struct CentroidWs {
    int* blockcnt;      // [nblocks][K]
    int* counts32;      // [K]
    long long* offsets; // [K+1]
    int* order;         // [n]
    float* partial;     // [K][SPLITS][dim]
};

static inline size_t centroid_ws_layout(long long n, int K, int dim, CentroidWs* w, char* base) {
    long long nblocks = ceil_div<long long>(n, CU_BLOCK_ROWS);
    size_t off = 0;
    auto take_ = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    size_t o_bc = take_((size_t)nblocks * K * 4);
    size_t o_c = take_((size_t)K * 4);
    size_t o_o = take_((size_t)(K + 1) * 8);
    size_t o_ord = take_((size_t)n * 4);
    size_t o_p = take_((size_t)K * CU_SPLITS * dim * 4);
    if (w) {
        w->blockcnt = (int*)(base + o_bc);
        w->counts32 = (int*)(base + o_c);
        w->offsets = (long long*)(base + o_o);
        w->order = (int*)(base + o_ord);
        w->partial = (float*)(base + o_p);
    }
    return off;
}

}  // namespace rqk

extern "C" {

size_t rqk_centroid_workspace_bytes(int64_t n, int32_t k, int32_t dim) {
    return rqk::centroid_ws_layout(n, k, dim, nullptr, nullptr);
}

// sums[k][dim] (fp32) and counts[k] (int64) of the rows of x assigned to each cluster.
// Deterministic.  Asynchronous on `stream`.
int rqk_centroid_accumulate(const float* x, int64_t n, int32_t dim, const int32_t* assign, int32_t k,
                            float* sums, int64_t* counts, void* workspace, size_t workspace_bytes,
                            void* stream_) {
    using namespace rqk;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !assign || !sums || !counts || !workspace) return fail(RQK_ERR_ARG, "rqk_centroid_accumulate: null pointer%s");
    if (k < 1 || k > 1024) return fail(RQK_ERR_UNSUPPORTED, "rqk_centroid_accumulate: k=%s%lld outside [1,1024]", "", k);
    if (dim % 4 != 0 || dim < 4 || dim > 4096) return fail(RQK_ERR_UNSUPPORTED, "rqk_centroid_accumulate: dim=%s%lld must be a multiple of 4 in [4,4096]", "", dim);
    if (n < 1 || n > 0x7fffffffLL) return fail(RQK_ERR_ARG, "rqk_centroid_accumulate: n=%s%lld outside [1,2^31)", "", n);
    CentroidWs w;
    size_t need = centroid_ws_layout(n, k, dim, &w, (char*)workspace);
    if (workspace_bytes < need) return fail(RQK_ERR_WORKSPACE, "rqk_centroid_accumulate: workspace %s%lld < %lld bytes", "", (long long)workspace_bytes, (long long)need);
    const int nblocks = (int)ceil_div<long long>(n, CU_BLOCK_ROWS);
    cu_block_count_kernel<<<nblocks, 1024, (size_t)k * 4, stream>>>(assign, n, k, w.blockcnt);
    cu_block_scan_kernel<<<k, 256, 0, stream>>>(w.blockcnt, nblocks, k, w.counts32);
    cu_offsets_kernel<<<1, 32, 0, stream>>>(w.counts32, k, w.offsets);
    size_t smem = (size_t)32 * k * 4;
    if (smem > 48 * 1024) RQK_CUDA_OK(cudaFuncSetAttribute(cu_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cu_scatter_kernel<<<nblocks, 1024, smem, stream>>>(assign, n, k, w.blockcnt, w.offsets, w.order);
    cu_partial_sum_kernel<<<dim3(k, CU_SPLITS), dim / 4, 0, stream>>>(x, dim, w.order, w.offsets, w.partial);
    cu_combine_kernel<<<k, 256, 0, stream>>>(w.partial, dim, sums);
    cu_counts_to_i64_kernel<<<ceil_div(k, 256), 256, 0, stream>>>(w.counts32, k, (long long*)counts);
    RQK_LAUNCH_OK();
    return 0;
}

// centers[k] <- sums[k]/counts[k] in place (centers holds the previous centroids on entry);
// shift_out[0] = sum_k |move_k|_2 (reference :343-346), shift_out[1] = #empty clusters, whose
// centres are left untouched and flagged in empty_mask[k] (the host redraws them, reference :321-322).
int rqk_centroid_finalize(const float* sums, const int64_t* counts, int32_t k, int32_t dim, float* centers,
                          float* shift_out, int32_t* empty_mask, void* stream_) {
    using namespace rqk;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!sums || !counts || !centers || !shift_out) return fail(RQK_ERR_ARG, "rqk_centroid_finalize: null pointer%s");
    if (k < 1 || k > 1024) return fail(RQK_ERR_UNSUPPORTED, "rqk_centroid_finalize: k=%s%lld outside [1,1024]", "", k);
    cu_finalize_kernel<<<1, 1024, 0, stream>>>(sums, (const long long*)counts, k, dim, centers, shift_out, empty_mask);
    RQK_LAUNCH_OK();
    return 0;
}

}  // extern "C"
