// Per-level residual of the hierarchy (hierarchical_rq_kmeans.py:1088-1128):
//   r = x - C[id];  for every dim-group g:  r_g /= (|r_g|_2 + 1e-8)
// and the per-dim weighting of _apply_weights (:583-604).  One warp per row, the row lives in
// registers between the two sweeps, so x is read once and r written once (8*D bytes per vector;
// in place when out == x).
#include "common.cuh"

namespace rqk {

constexpr int RS_MAX_PER_LANE = 32;   // dim <= 1024... (32 floats per lane)

// group_end[g] = exclusive end dim of group g (ascending), ngroups <= 32
__global__ void __launch_bounds__(256)
residual_kernel(const float* __restrict__ x, long long n, int dim, const int* __restrict__ ids,
                const float* __restrict__ centers, const int* __restrict__ group_end, int ngroups,
                float* __restrict__ out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* xr = x + row * dim;
    const float* cr = centers + (long long)ids[row] * dim;
    float r[RS_MAX_PER_LANE];
    const int nv = dim / 128;            // float4 per lane (dim % 128 == 0 fast path) else scalar path
    if (dim % 128 == 0 && ngroups == 1) {
        float ss = 0.f;
#pragma unroll
        for (int v = 0; v < RS_MAX_PER_LANE / 4; ++v) {
            if (v < nv) {
                float4 a = *reinterpret_cast<const float4*>(xr + v * 128 + lane * 4);
                float4 b = *reinterpret_cast<const float4*>(cr + v * 128 + lane * 4);
                r[v * 4 + 0] = a.x - b.x; r[v * 4 + 1] = a.y - b.y; r[v * 4 + 2] = a.z - b.z; r[v * 4 + 3] = a.w - b.w;
                ss = fmaf(r[v * 4 + 0], r[v * 4 + 0], ss); ss = fmaf(r[v * 4 + 1], r[v * 4 + 1], ss);
                ss = fmaf(r[v * 4 + 2], r[v * 4 + 2], ss); ss = fmaf(r[v * 4 + 3], r[v * 4 + 3], ss);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float den = sqrtf(ss) + 1e-8f;
        float* orow = out + row * dim;
#pragma unroll
        for (int v = 0; v < RS_MAX_PER_LANE / 4; ++v) {
            if (v < nv) {
                float4 o4 = make_float4(r[v * 4 + 0] / den, r[v * 4 + 1] / den, r[v * 4 + 2] / den, r[v * 4 + 3] / den);
                *reinterpret_cast<float4*>(orow + v * 128 + lane * 4) = o4;
            }
        }
        return;
    }
    // general path: arbitrary dim (<= 1024) and dim-groups; lane owns dims lane, lane+32, ...
    const int per = (dim + 31) / 32;
    for (int g = 0, start = 0; g < ngroups; ++g) {
        const int end = group_end[g];
        float ss = 0.f;
        for (int i = 0; i < per; ++i) {
            int d = lane + 32 * i;
            if (d >= start && d < end) {
                float v = xr[d] - cr[d];
                r[i] = v;
                ss = fmaf(v, v, ss);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float den = sqrtf(ss) + 1e-8f;
        for (int i = 0; i < per; ++i) {
            int d = lane + 32 * i;
            if (d >= start && d < end) out[row * dim + d] = r[i] / den;
        }
        start = end;
    }
}

__global__ void scale_dims_kernel(const float* __restrict__ x, long long total, int dim,
                                  const float* __restrict__ w, float* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < total; i += stride) out[i] = x[i] * w[i % dim];
}

__global__ void gather_rows_kernel(const float* __restrict__ x, int dim, const long long* __restrict__ rows,
                                   int nrows, float* __restrict__ out) {
    int r = blockIdx.x;
    if (r >= nrows) return;
    const float* src = x + rows[r] * dim;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) out[(long long)r * dim + d] = src[d];
}

// r = x - C[id] without normalisation (simplified_semantic_id_generator.py:78-96, :160-164): one warp per row.
__global__ void __launch_bounds__(256)
residual_plain_kernel(const float* __restrict__ x, long long n, int dim, const int* __restrict__ ids,
                      const float* __restrict__ centers, float* __restrict__ out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* xr = x + row * dim;
    const float* cr = centers + (long long)ids[row] * dim;
    float* orow = out + row * dim;
    if (dim % 4 == 0) {
        for (int d = lane * 4; d < dim; d += 128) {
            const float4 a = *reinterpret_cast<const float4*>(xr + d), b = *reinterpret_cast<const float4*>(cr + d);
            *reinterpret_cast<float4*>(orow + d) = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        }
    } else {
        for (int d = lane; d < dim; d += 32) orow[d] = xr[d] - cr[d];
    }
}

// ids[row] = argmin over the candidates c with allow[group[row]][c] != 0 of dist[row][c]; first index on ties, 0 if no
// candidate is allowed (torch.argmin of an all-inf row).  simplified_semantic_id_generator.py:317-331:
// `dist[batch_match_matrix == 0] = inf; argmin(dist, dim=1)`.  One warp per row.
// penalty == 0: a disallowed candidate counts as +inf (the Simplified generator); penalty != 0: as fl32(d + 10000),
// the reference's `distance.add_(10000.0 * (1 - match))` (hierarchical_rq_kmeans.py:953, :1288), one fp32 rounding.
__global__ void __launch_bounds__(256)
masked_argmin_kernel(const float* __restrict__ dist, long long n, int k, const int* __restrict__ group,
                     const unsigned char* __restrict__ allow, int penalty, int* __restrict__ out) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* dr = dist + row * k;
    const unsigned char* ar = allow + (long long)group[row] * k;
    float best = __int_as_float(0x7f800000);
    int arg = 0x7fffffff;
    for (int c = lane; c < k; c += 32) {
        const float d = ar[c] ? dr[c] : (penalty ? __fadd_rn(dr[c], 10000.0f) : __int_as_float(0x7f800000));
        if (d < best) { best = d; arg = c; }                     // strict: the first index wins inside a lane
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) out[row] = (arg == 0x7fffffff) ? 0 : arg;
}

int residual_launch(const float* x, long long n, int dim, const int* ids, const float* centers,
                    const int* group_end, int ngroups, float* out, cudaStream_t stream) {
    if (n == 0) return 0;
    residual_kernel<<<(unsigned)ceil_div<long long>(n * 32, 256), 256, 0, stream>>>(x, n, dim, ids, centers, group_end,
                                                                                   ngroups, out);
    RQK_LAUNCH_OK();
    return 0;
}

int scale_dims_launch(const float* x, long long n, int dim, const float* w, float* out, cudaStream_t stream) {
    if (n == 0) return 0;
    scale_dims_kernel<<<148 * 8, 256, 0, stream>>>(x, n * dim, dim, w, out);
    RQK_LAUNCH_OK();
    return 0;
}

}  // namespace rqk

extern "C" {

// out[n][dim] = group-normalised (x - centers[ids]); out may alias x.  group_end: DEVICE int32[ngroups].
int rqk_residual_normalise(const float* x, int64_t n, int32_t dim, const int32_t* ids, const float* centers,
                           const int32_t* group_end, int32_t ngroups, float* out, void* stream_) {
    using namespace rqk;
    if (!x || !ids || !centers || !out || !group_end) return fail(RQK_ERR_ARG, "rqk_residual_normalise: null pointer%s");
    if (dim < 1 || dim > 1024) return fail(RQK_ERR_UNSUPPORTED, "rqk_residual_normalise: dim=%s%lld outside [1,1024]", "", dim);
    if (ngroups < 1 || ngroups > 32) return fail(RQK_ERR_UNSUPPORTED, "rqk_residual_normalise: ngroups=%s%lld outside [1,32]", "", ngroups);
    return residual_launch(x, n, dim, ids, centers, group_end, ngroups, out, (cudaStream_t)stream_);
}

// out[n][dim] = x - centers[ids] (no normalisation: simplified_semantic_id_generator.py:78-96); out may alias x.
int rqk_residual_plain(const float* x, int64_t n, int32_t dim, const int32_t* ids, const float* centers, float* out,
                       void* stream_) {
    using namespace rqk;
    if (!x || !ids || !centers || !out) return fail(RQK_ERR_ARG, "rqk_residual_plain: null pointer%s");
    if (dim < 1) return fail(RQK_ERR_ARG, "rqk_residual_plain: dim=%s%lld < 1", "", dim);
    if (n == 0) return 0;
    residual_plain_kernel<<<(unsigned)ceil_div<long long>(n * 32, 256), 256, 0, (cudaStream_t)stream_>>>(x, n, dim, ids, centers, out);
    RQK_LAUNCH_OK();
    return 0;
}

// ids[n] = first argmin of dist[n][k] over the candidates allowed for the row's group (allow: uint8 [ngroups][k]);
// penalty 0: disallowed = +inf; 1: disallowed = fl32(d + 10000) (the reference's last-layer match-matrix mask).
int rqk_masked_argmin(const float* dist, int64_t n, int32_t k, const int32_t* group, const uint8_t* allow,
                      int32_t ngroups, int32_t penalty, int32_t* ids, void* stream_) {
    using namespace rqk;
    if (!dist || !group || !allow || !ids) return fail(RQK_ERR_ARG, "rqk_masked_argmin: null pointer%s");
    if (k < 1 || ngroups < 1) return fail(RQK_ERR_ARG, "rqk_masked_argmin: k=%s%lld, ngroups=%lld must be >= 1", "", k, ngroups);
    if (n == 0) return 0;
    masked_argmin_kernel<<<(unsigned)ceil_div<long long>(n * 32, 256), 256, 0, (cudaStream_t)stream_>>>(dist, n, k, group, allow, penalty, ids);
    RQK_LAUNCH_OK();
    return 0;
}

// out = x * w (per-dim weights, hierarchical_rq_kmeans.py:583-604); out may alias x.
int rqk_scale_dims(const float* x, int64_t n, int32_t dim, const float* w, float* out, void* stream_) {
    using namespace rqk;
    if (!x || !w || !out) return fail(RQK_ERR_ARG, "rqk_scale_dims: null pointer%s");
    return scale_dims_launch(x, n, dim, w, out, (cudaStream_t)stream_);
}

// out[r] = x[rows[r]] : centroid (re-)initialisation from host-drawn row indices
// (balancekmeans/__init__.py:240-256).  rows: DEVICE int64[nrows].
int rqk_gather_rows(const float* x, int32_t dim, const int64_t* rows, int32_t nrows, float* out, void* stream_) {
    using namespace rqk;
    if (!x || !rows || !out) return fail(RQK_ERR_ARG, "rqk_gather_rows: null pointer%s");
    if (nrows == 0) return 0;
    gather_rows_kernel<<<nrows, 128, 0, (cudaStream_t)stream_>>>(x, dim, (const long long*)rows, nrows, out);
    RQK_LAUNCH_OK();
    return 0;
}

}  // extern "C"
