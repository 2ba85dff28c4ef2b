// Shared helpers for the rqk sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace rqk {

// ---- error plumbing: no exceptions cross the C ABI; message via rqk_last_error() ----
extern thread_local char g_last_error[512];

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
    snprintf(g_last_error, sizeof(g_last_error), fmt, a, b, c);
    return code;
}

#define RQK_ERR_ARG (-1)
#define RQK_ERR_CUDA (-2)
#define RQK_ERR_WORKSPACE (-3)
#define RQK_ERR_UNSUPPORTED (-4)
#define RQK_ERR_INTERNAL (-5)

#define RQK_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return rqk::fail(RQK_ERR_CUDA, "%s: CUDA error %lld at line %lld", cudaGetErrorString(_e), \
                             (long long)_e, (long long)__LINE__);                              \
    } while (0)

#define RQK_LAUNCH_OK()                                                                        \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return rqk::fail(RQK_ERR_CUDA, "%s: kernel launch failed (%lld) at line %lld",     \
                             cudaGetErrorString(_e), (long long)_e, (long long)__LINE__);      \
    } while (0)

// ---- fp16 bit helpers (device) ----
// Monotone map of fp16 bit patterns onto uint16 so that unsigned order == numeric order.
// -0 is folded onto +0 (torch compares them equal); NaNs are excluded by contract
// (reference asserts none: balancekmeans/__init__.py:36).
__device__ __forceinline__ uint32_t h2key(uint32_t h) {
    h = (h == 0x8000u) ? 0u : h;
    return (h & 0x8000u) ? (~h & 0xffffu) : (h | 0x8000u);
}
__device__ __forceinline__ uint32_t key2h(uint32_t k) {
    return (k & 0x8000u) ? (k & 0x7fffu) : (~k & 0xffffu);
}
__device__ __forceinline__ __half bits2h(uint32_t b) { return __ushort_as_half((unsigned short)b); }
__device__ __forceinline__ uint32_t h2bits(__half h) { return (uint32_t)__half_as_ushort(h); }

template <typename T>
__host__ __device__ __forceinline__ T ceil_div(T a, T b) { return (a + b - 1) / b; }

template <typename T>
__host__ __device__ __forceinline__ T round_up(T a, T b) { return ceil_div(a, b) * b; }

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// per-device caches (function attributes, events) are indexed by the CUDA device ordinal & (RQK_MAX_DEVICES - 1)
constexpr int RQK_MAX_DEVICES = 16;

// Everything the score pass can emit (any pointer may be null), shared by the tcgen05 kernel and
// the CUDA-core cross-check kernel.
struct ScoreOut {
    __half* scores_t;            // [k][ld] fp16 = half(-dist), worker-major, columns >= n hold -inf
    long long ld;
    int* argmin;                 // [n] first index of the smallest distance
    float* best2;                // [n][2] {smallest, second smallest} distance
    int* counts;                 // [k] += bincount(argmin)
    unsigned int* minmax_keys;   // [2] {max key, min key} of scores_t (monotone fp16 keys)
    int farthest;                // 1: argmax of the distance instead (N<K quirk, balancekmeans/__init__.py:24-26)
    // predict()-mode masking (hierarchical_rq_kmeans.py:1210-1219): centres outside
    // [mask_ids[n]*mask_block, (mask_ids[n]+1)*mask_block) get fl32(d + 10000) before the argmin
    const int* mask_ids;
    int mask_block;
    float* dist;                 // [n][k] fp32 distances, row-major (API parity with pairwise_distance_full only)
};

int score_pass_simt(const float* x, long long n, int dim, const float* c, int K, float* x2, float* c2,
                    const ScoreOut& o, cudaStream_t stream);
int score_pass_tc(const float* x, long long n, int dim, const float* c, int K, float* chi, float* clo, float* c2,
                  const ScoreOut& o, int num_sms, cudaStream_t stream);
int score_pass_dispatch(const float* x, long long n, int dim, const float* c, int K, const ScoreOut& o, int flags,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t score_workspace_bytes(long long n, int K, int dim);
int residual_launch(const float* x, long long n, int dim, const int* ids, const float* centers,
                    const int* group_end, int ngroups, float* out, cudaStream_t stream);
int scale_dims_launch(const float* x, long long n, int dim, const float* w, float* out, cudaStream_t stream);

}  // namespace rqk
