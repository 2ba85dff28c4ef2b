// Multi-level encode in ONE tcgen05 kernel: ids of all levels from one HBM read of X, and the per-level residual
//   r_{l+1} = (r_l - C_l[id_l]) / (|r_l - C_l[id_l]|_2 + 1e-8)      (hierarchical_rq_kmeans.py:1111-1122)
// is never written anywhere - not to HBM, not to shared memory.  (Reference call chain: predict :539-581 over
// _predict_layer_0/_middle/_last :1146-1305, and the id chain train() emits :654-660.)
//
// The residual is linear in x, so every score a later level needs is a correction of a dot product with the
// ORIGINAL row:
//     r_{l+1} . v = (r_l . v - C_l[id_l] . v) / s_l,              s_l = |r_l - C_l[id_l]| + 1e-8 = d_l + 1e-8
//     r_m . C_m[k] = (..((x . C_m[k] - G_0m[id_0][k]) / s_0 - G_1m[id_1][k]) / s_1 ..) / s_{m-1},   G_lm = C_l C_m^T
//     |r_{l+1}|^2  = (d_l / s_l)^2
// and d_l, the distance to the chosen centre, is what the level's argmin has just produced.  So the contraction is
// X . [C_0; C_1; ...; C_{L-1}]^T on the tensor cores (3xTF32, exactly the pipeline of score_tc.cu), and the epilogue
// thread that owns a row walks the levels: argmin of level 0, then the level-1 scores from the level-1 columns and one
// gathered row of the K_0 x K_1 table G_01, and so on.  [128,128,256] is 512 accumulator columns = all of TMEM, so the
// columns are split into PHASES of at most 256 (levels 0+1 | level 2): a CTA runs the k-loop of a 128-row tile once per
// phase into one of two 256-column accumulators, and the epilogue of a phase overlaps the MMAs of the next one (the
// second pass over the tile's X blocks is served by L2: 256 KB per tile, re-read microseconds later).  The row state
// (ids, scales) lives in the registers of the row's epilogue thread across the phases.
//
// Accuracy.  The corrected scores carry the tensor cores' accumulation error of the base dot products divided by the
// product of the scales (~1.2x per level), i.e. they are slightly less accurate than a score pass over a materialised
// residual.  Rows whose top-2 gap at some level is inside that error budget are FLAGGED (a few per 100 000) and a small
// second kernel re-evaluates them from scratch through the literal chain in fp32 (residual vector in shared memory,
// full distance scan per level), so a flagged row's ids do not depend on the algebra at all.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace rqk {

constexpr int EF_MAX_LEVELS = 4;
constexpr int EF_MAX_PHASES = 4;
constexpr int EF_MAX_KSUM = 1024;      // sum of the levels' cluster counts (|c|^2 table in shared memory)

struct EncFusedParams {
    long long n;
    int dim, levels, phases;
    int wmax;                          // centroid rows per phase block of the concatenated hi / lo matrices (= TMA box rows)
    int tmem_cols;                     // columns per accumulator stage (power of two >= wmax)
    int stages, raw;
    int K[EF_MAX_LEVELS];
    int col_of[EF_MAX_LEVELS];         // first accumulator column of the level inside its phase
    int c2_of[EF_MAX_LEVELS];          // offset of the level in the |c|^2 table
    int width[EF_MAX_PHASES];          // UMMA N of the phase (multiple of 16)
    int first_level[EF_MAX_PHASES + 1];
    int masked[EF_MAX_LEVELS];         // predict() mode: level adds fl32(+10000) outside the previous id's block
    int mask_block[EF_MAX_LEVELS];
    const float* c2;                   // [sum K] |c|^2
    const float* G[EF_MAX_LEVELS][EF_MAX_LEVELS];   // G[m][l], m < l: [K_m][K_l] fp32 = C_m C_l^T
    int* ids;                          // [levels][n]
    float flag_tol;                    // multiple of the error budget below which a top-2 gap flags the row
    int* flag_list;                    // [n] rows to re-evaluate
    int* flag_count;
};

// One level of one row.  LV is a compile-time level index so that the id / scale arrays stay in registers.
// Returns the winner (bi), its d^2 (wd2) and whether the decision is inside the error budget.
template <int LV>
__device__ __forceinline__ void ef_level(const EncFusedParams& P, uint32_t taddr, const float* __restrict__ c2s,
                                         float rn2, float err, const int (&id)[EF_MAX_LEVELS],
                                         const float (&inv_s)[EF_MAX_LEVELS], int& bi, float& wd2, bool& risky) {
    const int K = P.K[LV];
    const float* cc = c2s + P.c2_of[LV];
    // predict(): every centre outside [pid*block, (pid+1)*block) gets fl32(d + 10000) (hierarchical_rq_kmeans.py:1210-1219)
    int mlo = 0, mhi = K;
    bool anymask = false;
    if (LV > 0 && P.masked[LV]) {
        mlo = id[LV > 0 ? LV - 1 : 0] * P.mask_block[LV];
        mhi = mlo + P.mask_block[LV];
        anymask = (mlo > 0) || (mhi < K);
    }
    const float* grow[LV > 0 ? LV : 1];
#pragma unroll
    for (int m = 0; m < LV; ++m) grow[m] = P.G[m][LV] + (size_t)id[m] * K;
    float b1 = INFINITY, b2 = INFINITY;
    bi = 0;
    wd2 = 0.f;
    for (int c0 = 0; c0 < K; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float g[LV > 0 ? LV : 1][16];
#pragma unroll
            for (int m = 0; m < LV; ++m) {
                const float4* gp = reinterpret_cast<const float4*>(grow[m] + c0 + h * 16);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float4 q4 = __ldg(gp + t);
                    g[m][t * 4 + 0] = q4.x; g[m][t * 4 + 1] = q4.y; g[m][t * 4 + 2] = q4.z; g[m][t * 4 + 3] = q4.w;
                }
            }
            if (h == 0) tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = c0 + h * 16 + j;
                float dot = __uint_as_float(v[h * 16 + j]);
#pragma unroll
                for (int m = 0; m < LV; ++m) dot = (dot - g[m][j]) * inv_s[m];
                const float d2 = fmaxf(fmaf(-2.f, dot, rn2 + cc[k]), 0.f);
                float val = d2;
                if (anymask) {
                    val = sqrtf(d2);
                    if (k < mlo || k >= mhi) val = val + 10000.0f;
                }
                if (val < b1) { b2 = b1; b1 = val; bi = k; wd2 = d2; }
                else if (val < b2) b2 = val;
            }
        }
    }
    // unmasked decision: compared on d^2, flagged when the runner-up is inside the error budget.  Masked rows are
    // decided on the fl32(d + 10000) grid (first index among equal values); no flagging there.
    risky = !anymask && (b2 - b1) < P.flag_tol * err;
}

template <int LV>
__device__ __forceinline__ void ef_level_step(const EncFusedParams& P, int lo, int hi, uint32_t tacc,
                                              const float* __restrict__ c2s, float xn0, float& rn2, float& amp,
                                              float& rel, int (&id)[EF_MAX_LEVELS], float (&inv_s)[EF_MAX_LEVELS],
                                              bool& flagged) {
    if (LV < lo || LV >= hi) return;
    int bi; float wd2; bool risky;
    // error budget of this level's d^2: the base dot products' error (measured bound 7e-6 (|x|^2 + |c|^2)) divided by the
    // scales so far, plus what the relative error of those scales does to a dot product of size ~|c|
    const float c2r = c2s[P.c2_of[LV]];
    const float err = 7e-6f * (xn0 + c2r) * amp + 2.f * rel * sqrtf(c2r);
    ef_level<LV>(P, tacc + (uint32_t)P.col_of[LV], c2s, rn2, err, id, inv_s, bi, wd2, risky);
    flagged |= risky;
    id[LV] = bi;
    const float d = sqrtf(wd2);
    const float s = d + 1e-8f;
    inv_s[LV] = 1.0f / s;
    rel += err / fmaxf(2.f * wd2, 1e-30f);     // s^2 = d^2: relative error of the scale
    amp *= inv_s[LV];
    const float ratio = d * inv_s[LV];
    rn2 = ratio * ratio;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
encode_fused_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_chi,
                    const __grid_constant__ CUtensorMap map_clo, const __grid_constant__ EncFusedParams P) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = P.stages, nraw = P.raw, phases = P.phases;
    const int kblocks = P.dim / TC_BK;
    const long long ntiles = (P.n + TC_BM - 1) / TC_BM;

    // ---- smem carve-up (as score_tc.cu): per stage {A hi 16K, A lo 16K, B hi wmax*128, B lo wmax*128}, raw X ring ----
    const uint32_t a_bytes = TC_BM * TC_BK * 4;
    const uint32_t b_bytes = (uint32_t)P.wmax * TC_BK * 4;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    unsigned char* raw0 = smem + (size_t)stages * stage_bytes;
    unsigned char* tail = raw0 + (size_t)nraw * a_bytes;
    uint64_t* full_bar = (uint64_t*)tail;          // [stages]
    uint64_t* xf_bar = full_bar + 8;               // [stages]
    uint64_t* empty_bar = xf_bar + 8;              // [stages]
    uint64_t* tmem_full = empty_bar + 8;           // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint64_t* x2_full = tmem_empty + 2;            // [2]
    uint64_t* raw_full = x2_full + 2;              // [TC_MAX_RAW]
    uint64_t* raw_empty = raw_full + TC_MAX_RAW;   // [TC_MAX_RAW]
    uint32_t* tmem_base_slot = (uint32_t*)(raw_empty + TC_MAX_RAW);
    float* x2s = (float*)(tmem_base_slot + 4);     // [2][128]
    float* c2s = x2s + 2 * TC_BM;                  // [EF_MAX_KSUM]

    int ksum = 0;
    for (int l = 0; l < P.levels; ++l) ksum += P.K[l];
    for (int i = threadIdx.x; i < ksum; i += TC_THREADS) c2s[i] = P.c2[i];
    if (warp == TC_TMA_WARP && lane == 0) {
        tma_prefetch_desc(&map_x);
        tma_prefetch_desc(&map_chi);
        tma_prefetch_desc(&map_clo);
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&xf_bar[s], 128);
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < nraw; ++i) {
            mbar_init(&raw_full[i], 1);
            mbar_init(&raw_empty[i], 128);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 128);
            mbar_init(&x2_full[a], 128);
        }
        fence_barrier_init();
    }
    if (warp == TC_MMA_WARP) {
        uint32_t ncols = 2u * (uint32_t)P.tmem_cols;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == TC_XLOAD_WARP) {
        // ===================== TMA producer: X blocks into the raw ring, once per phase =====================
        if (lane == 0) {
            int rs = 0; uint32_t rph = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int row0 = (int)(tile * TC_BM);
                for (int p = 0; p < phases; ++p) {
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&raw_empty[rs], rph ^ 1);
                        mbar_expect_tx(&raw_full[rs], a_bytes);
                        tma_load_2d(raw0 + (size_t)rs * a_bytes, &map_x, &raw_full[rs], kb * TC_BK, row0);
                        if (++rs == nraw) { rs = 0; rph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == TC_TMA_WARP) {
        // ===================== TMA producer: the phase's centroid hi / lo blocks =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int p = 0; p < phases; ++p) {
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        unsigned char* st = smem + (size_t)s * stage_bytes;
                        mbar_expect_tx(&full_bar[s], 2 * b_bytes);
                        tma_load_2d(st + 2 * a_bytes, &map_chi, &full_bar[s], kb * TC_BK, p * P.wmax);
                        tma_load_2d(st + 2 * a_bytes + b_bytes, &map_clo, &full_bar[s], kb * TC_BK, p * P.wmax);
                        if (++s == stages) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == TC_MMA_WARP) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            int a = 0; uint32_t aph = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int p = 0; p < phases; ++p) {
                    const uint32_t idesc = make_idesc_tf32(TC_BM, P.width[p]);
                    mbar_wait(&tmem_empty[a], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(a * P.tmem_cols);
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&full_bar[s], ph);
                        mbar_wait(&xf_bar[s], ph);
                        tc_fence_after();
                        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                        const uint64_t a_hi = make_sw128_desc(st);
                        const uint64_t a_lo = make_sw128_desc(st + a_bytes);
                        const uint64_t b_hi = make_sw128_desc(st + 2 * a_bytes);
                        const uint64_t b_lo = make_sw128_desc(st + 2 * a_bytes + b_bytes);
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) {
                            const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);
                            umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
                            umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                            umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
                        }
                        umma_commit(&empty_bar[s]);
                        if (kb == kblocks - 1) umma_commit(&tmem_full[a]);
                        if (++s == stages) { s = 0; ph ^= 1; }
                    }
                    if (++a == 2) { a = 0; aph ^= 1; }
                }
            }
        }
    } else if (warp >= TC_XF_WARP0 && warp < TC_XF_WARP0 + 4) {
        // ===================== transform: hi/lo split + row norms (identical to score_tc.cu) =====================
        const int r = (warp - TC_XF_WARP0) * 32 + lane;
        int s = 0; uint32_t ph = 0;
        int rs = 0; uint32_t rph = 0;
        int a = 0; uint32_t aph = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int p = 0; p < phases; ++p) {
                float nrm = 0.f;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&raw_full[rs], rph);
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    unsigned char* st = smem + (size_t)s * stage_bytes;
                    const float4* raw = reinterpret_cast<const float4*>(raw0 + (size_t)rs * a_bytes + (size_t)r * 128);
                    float4* hi = reinterpret_cast<float4*>(st + (size_t)r * 128);
                    float4* lo = reinterpret_cast<float4*>(st + a_bytes + (size_t)r * 128);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int ch = c ^ (lane & 7);
                        float4 v = raw[ch];
                        nrm = fmaf(v.x, v.x, nrm); nrm = fmaf(v.y, v.y, nrm);
                        nrm = fmaf(v.z, v.z, nrm); nrm = fmaf(v.w, v.w, nrm);
                        float4 h, l;
                        h.x = to_tf32_rna(v.x); h.y = to_tf32_rna(v.y); h.z = to_tf32_rna(v.z); h.w = to_tf32_rna(v.w);
                        l.x = to_tf32_rna(v.x - h.x); l.y = to_tf32_rna(v.y - h.y);
                        l.z = to_tf32_rna(v.z - h.z); l.w = to_tf32_rna(v.w - h.w);
                        hi[ch] = h;
                        lo[ch] = l;
                    }
                    fence_proxy_async();
                    mbar_arrive(&xf_bar[s]);
                    mbar_arrive(&raw_empty[rs]);
                    if (++s == stages) { s = 0; ph ^= 1; }
                    if (++rs == nraw) { rs = 0; rph ^= 1; }
                }
                mbar_wait(&tmem_empty[a], aph ^ 1);
                x2s[a * TC_BM + r] = nrm;
                mbar_arrive(&x2_full[a]);
                if (++a == 2) { a = 0; aph ^= 1; }
            }
        }
    } else if (warp < 4) {
        // ===================== epilogue: the row's thread walks the levels =====================
        const int q = warp;
        const int r = q * 32 + lane;
        int a = 0; uint32_t aph = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const long long row = tile * TC_BM + r;
            const bool rv = row < P.n;
            int id[EF_MAX_LEVELS] = {0, 0, 0, 0};
            float inv_s[EF_MAX_LEVELS] = {1.f, 1.f, 1.f, 1.f};
            float rn2 = 0.f, xn0 = 0.f, amp = 1.f, rel = 0.f;
            bool flagged = false;
            for (int p = 0; p < phases; ++p) {
                mbar_wait(&x2_full[a], aph);
                if (p == 0) { xn0 = x2s[a * TC_BM + r]; rn2 = xn0; }
                mbar_wait(&tmem_full[a], aph);
                tc_fence_after();
                const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * P.tmem_cols);
                const int lo = P.first_level[p], hi = P.first_level[p + 1];
                ef_level_step<0>(P, lo, hi, tacc, c2s, xn0, rn2, amp, rel, id, inv_s, flagged);
                ef_level_step<1>(P, lo, hi, tacc, c2s, xn0, rn2, amp, rel, id, inv_s, flagged);
                ef_level_step<2>(P, lo, hi, tacc, c2s, xn0, rn2, amp, rel, id, inv_s, flagged);
                ef_level_step<3>(P, lo, hi, tacc, c2s, xn0, rn2, amp, rel, id, inv_s, flagged);
                tc_fence_before();
                mbar_arrive(&tmem_empty[a]);
                if (++a == 2) { a = 0; aph ^= 1; }
            }
            if (rv) {
#pragma unroll
                for (int l = 0; l < EF_MAX_LEVELS; ++l)
                    if (l < P.levels) P.ids[(long long)l * P.n + row] = id[l];
                if (flagged) P.flag_list[atomicAdd(P.flag_count, 1)] = (int)row;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == TC_MMA_WARP) {
        uint32_t ncols = 2u * (uint32_t)P.tmem_cols;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// ---- flagged rows: the literal chain, one CTA per row, residual vector in shared memory ----
struct EncFixParams {
    long long n;
    int dim, levels;
    int K[EF_MAX_LEVELS];
    int c2_of[EF_MAX_LEVELS];
    int masked[EF_MAX_LEVELS];
    int mask_block[EF_MAX_LEVELS];
    const float* C[EF_MAX_LEVELS];
    const float* c2;
    int* ids;
    const int* flag_list;
    const int* flag_count;
};

constexpr int EFX_THREADS = 256;

__device__ __forceinline__ float efx_block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                       // red[] from an earlier call has been consumed
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < EFX_THREADS / 32; ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(EFX_THREADS)
encode_fixup_kernel(const float* __restrict__ x, const EncFixParams F) {
    extern __shared__ float efx_sm[];
    float* rvec = efx_sm;                  // [dim]
    float* dist = rvec + F.dim;            // [256]
    float* red = dist + 256;               // [8]
    __shared__ int s_id;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int count = *F.flag_count;
    for (int e = blockIdx.x; e < count; e += gridDim.x) {
        const long long row = F.flag_list[e];
        __syncthreads();
        for (int d = threadIdx.x; d < F.dim; d += EFX_THREADS) rvec[d] = x[row * F.dim + d];
        __syncthreads();
        int prev = 0;
        for (int l = 0; l < F.levels; ++l) {
            const int K = F.K[l];
            const float* C = F.C[l];
            const float* c2 = F.c2 + F.c2_of[l];
            float ss = 0.f;
            for (int d = threadIdx.x; d < F.dim; d += EFX_THREADS) ss = fmaf(rvec[d], rvec[d], ss);
            const float rn2 = efx_block_sum(ss, red);
            int mlo = 0, mhi = K;
            if (l > 0 && F.masked[l]) { mlo = prev * F.mask_block[l]; mhi = mlo + F.mask_block[l]; }
            const bool anymask = (mlo > 0) || (mhi < K);
            for (int k = warp; k < K; k += EFX_THREADS / 32) {
                const float* ck = C + (size_t)k * F.dim;
                float dot = 0.f;
                for (int d = lane; d < F.dim; d += 32) dot = fmaf(rvec[d], ck[d], dot);
#pragma unroll
                for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (lane == 0) {
                    const float d2 = fmaxf(fmaf(-2.f, dot, rn2 + c2[k]), 0.f);
                    float val = d2;
                    if (anymask) {
                        val = sqrtf(d2);
                        if (k < mlo || k >= mhi) val = val + 10000.0f;
                    }
                    dist[k] = val;
                }
            }
            __syncthreads();
            if (warp == 0) {                                   // first index of the minimum
                float bv = INFINITY; int bk = 0;
                for (int k = lane; k < K; k += 32) {
                    const float v = dist[k];
                    if (v < bv) { bv = v; bk = k; }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                    if (ov < bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
                }
                if (lane == 0) { s_id = bk; F.ids[(long long)l * F.n + row] = bk; }
            }
            __syncthreads();
            prev = s_id;
            if (l < F.levels - 1) {
                const float* cw = C + (size_t)prev * F.dim;
                float s2 = 0.f;
                for (int d = threadIdx.x; d < F.dim; d += EFX_THREADS) {
                    const float v = rvec[d] - cw[d];
                    rvec[d] = v;                               // each thread re-reads only its own elements below
                    s2 = fmaf(v, v, s2);
                }
                const float den = sqrtf(efx_block_sum(s2, red)) + 1e-8f;
                for (int d = threadIdx.x; d < F.dim; d += EFX_THREADS) rvec[d] = rvec[d] / den;
                __syncthreads();
            }
        }
    }
}

// G[i][j] = C_a[i] . C_b[j], accumulated in fp64 (the tables are tiny; their error must not add to the budget)
__global__ void __launch_bounds__(128)
centroid_gram_kernel(const float* __restrict__ ca, int ka, const float* __restrict__ cb, int kb, int dim,
                     float* __restrict__ g) {
    const int i = blockIdx.y;
    const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= kb) return;
    double s = 0.0;
    for (int d = lane; d < dim; d += 32) s += (double)ca[(size_t)i * dim + d] * (double)cb[(size_t)j * dim + d];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) g[(size_t)i * kb + j] = (float)s;
}

__global__ void ef_clear_kernel(int* count) { *count = 0; }

struct EncFusedPlan {
    int levels, phases, wmax, ksum;
    int K[EF_MAX_LEVELS], col_of[EF_MAX_LEVELS], c2_of[EF_MAX_LEVELS], phase_of[EF_MAX_LEVELS];
    int width[EF_MAX_PHASES], first_level[EF_MAX_PHASES + 1];
};

// Packs consecutive levels into phases of at most 256 accumulator columns.  Returns false if the shape is outside
// what the kernel takes (the caller then uses the level chain, rqk_encode).
static bool ef_plan(int dim, int levels, const int32_t* ks, EncFusedPlan& pl) {
    if (levels < 1 || levels > EF_MAX_LEVELS || dim < 32 || dim % 32 != 0 || dim > 4096) return false;
    pl.levels = levels;
    pl.phases = 0;
    pl.ksum = 0;
    int w = 0;
    for (int l = 0; l < levels; ++l) {
        const int k = ks[l];
        if (k < 32 || k > 256 || k % 32 != 0) return false;
        if (pl.phases == 0 || w + k > 256) {
            if (pl.phases == EF_MAX_PHASES) return false;
            if (pl.phases) pl.width[pl.phases - 1] = w;
            pl.first_level[pl.phases++] = l;
            w = 0;
        }
        pl.K[l] = k;
        pl.phase_of[l] = pl.phases - 1;
        pl.col_of[l] = w;
        pl.c2_of[l] = pl.ksum;
        w += k;
        pl.ksum += k;
    }
    pl.width[pl.phases - 1] = w;
    pl.first_level[pl.phases] = levels;
    pl.wmax = 0;
    for (int p = 0; p < pl.phases; ++p) pl.wmax = pl.width[p] > pl.wmax ? pl.width[p] : pl.wmax;
    return pl.ksum <= EF_MAX_KSUM;
}

struct EncFusedWs {
    float *chi, *clo, *c2, *g[EF_MAX_LEVELS][EF_MAX_LEVELS];
    int *count, *list;
    size_t bytes;
};

static void ef_carve(const EncFusedPlan& pl, long long n, int dim, char* base, EncFusedWs& w) {
    char* p = base;
    w.count = (int*)p; p += 256;
    const size_t cb = align256((size_t)pl.phases * pl.wmax * dim * 4);
    w.chi = (float*)p; p += cb;
    w.clo = (float*)p; p += cb;
    w.c2 = (float*)p; p += align256((size_t)pl.ksum * 4);
    for (int m = 0; m < pl.levels; ++m)
        for (int l = m + 1; l < pl.levels; ++l) {
            w.g[m][l] = (float*)p;
            p += align256((size_t)pl.K[m] * pl.K[l] * 4);
        }
    w.list = (int*)p; p += align256((size_t)(n > 0 ? n : 1) * 4);
    w.bytes = (size_t)(p - base);
}

}  // namespace rqk

extern "C" {

// 1 if rqk_encode_fused takes this shape (1..4 levels, every cluster count a multiple of 32 in [32,256], dim a multiple
// of 32; predict() mode additionally needs needs[l] == ks[l] at the masked levels), else 0.
int rqk_encode_fused_supported(int32_t dim, int32_t levels, const int32_t* ks, const int32_t* needs, int32_t mode) {
    using namespace rqk;
    EncFusedPlan pl;
    if (!ks || !ef_plan(dim, levels, ks, pl)) return 0;
    if (mode == 1) {
        if (!needs) return 0;
        for (int l = 1; l < levels - 1; ++l)
            if (needs[l] != ks[l]) return 0;
    } else if (mode != 0) {
        return 0;
    }
    return 1;
}

size_t rqk_encode_fused_workspace_bytes(int64_t n, int32_t dim, int32_t levels, const int32_t* ks) {
    using namespace rqk;
    EncFusedPlan pl;
    if (!ks || !ef_plan(dim, levels, ks, pl)) return 0;
    EncFusedWs w;
    ef_carve(pl, n, dim, nullptr, w);
    return w.bytes + 256;
}

// Multi-level ids of x [n][dim] (unit weights, one dim-group) in one tensor-core kernel plus the re-evaluation of
// the flagged rows.  centers: HOST array [levels] of DEVICE pointers; ks / needs: HOST int32 [levels];
// ids: DEVICE int32 [levels][n]; mode as rqk_encode.  The first int32 of the workspace holds the number of flagged
// rows of the call once the stream has drained.
int rqk_encode_fused(const float* x, int64_t n, int32_t dim, int32_t levels, const void* const* centers,
                     const int32_t* ks, const int32_t* needs, int32_t* ids, int32_t mode, void* workspace,
                     size_t workspace_bytes, void* stream_) {
    using namespace rqk;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !centers || !ks || !needs || !ids || !workspace) return fail(RQK_ERR_ARG, "rqk_encode_fused: null pointer%s");
    if (n < 0 || n > 0x7fffffffLL) return fail(RQK_ERR_ARG, "rqk_encode_fused: n=%s%lld outside [0,2^31)", "", n);
    if (!rqk_encode_fused_supported(dim, levels, ks, needs, mode))
        return fail(RQK_ERR_UNSUPPORTED, "rqk_encode_fused: shape not supported (levels=%s%lld); use rqk_encode", "", levels);
    if (((uintptr_t)x & 15) || ((uintptr_t)workspace & 255)) return fail(RQK_ERR_ARG, "rqk_encode_fused: x must be 16-byte and the workspace 256-byte aligned%s");
    EncFusedPlan pl;
    ef_plan(dim, levels, ks, pl);
    EncFusedWs w;
    ef_carve(pl, n, dim, (char*)workspace, w);
    if (workspace_bytes < w.bytes)
        return fail(RQK_ERR_WORKSPACE, "rqk_encode_fused: workspace %s%lld < %lld bytes", "", (long long)workspace_bytes, (long long)w.bytes);
    ef_clear_kernel<<<1, 1, 0, stream>>>(w.count);
    if (n == 0) return 0;

    // concatenated hi / lo centroid matrices (phase p = rows [p*wmax, p*wmax + width[p]), zero padding), |c|^2, Gram tables
    const size_t cb = (size_t)pl.phases * pl.wmax * dim * 4;
    RQK_CUDA_OK(cudaMemsetAsync(w.chi, 0, cb, stream));
    RQK_CUDA_OK(cudaMemsetAsync(w.clo, 0, cb, stream));
    for (int l = 0; l < levels; ++l) {
        if (!centers[l] || ((uintptr_t)centers[l] & 15)) return fail(RQK_ERR_ARG, "rqk_encode_fused: centers[%s%lld] null or unaligned", "", l);
        const size_t row0 = (size_t)pl.phase_of[l] * pl.wmax + pl.col_of[l];
        centroid_split_kernel<<<pl.K[l], 128, 0, stream>>>((const float*)centers[l], pl.K[l], dim, w.chi + row0 * dim,
                                                           w.clo + row0 * dim, w.c2 + pl.c2_of[l]);
        for (int m = 0; m < l; ++m)
            centroid_gram_kernel<<<dim3((unsigned)ceil_div(pl.K[l], 4), (unsigned)pl.K[m]), 128, 0, stream>>>(
                (const float*)centers[m], pl.K[m], (const float*)centers[l], pl.K[l], dim, w.g[m][l]);
    }
    RQK_LAUNCH_OK();

    EncFusedParams P;
    memset(&P, 0, sizeof(P));
    P.n = n; P.dim = dim; P.levels = levels; P.phases = pl.phases; P.wmax = pl.wmax;
    int tcols = 32;
    while (tcols < pl.wmax) tcols <<= 1;
    P.tmem_cols = tcols;
    for (int l = 0; l < levels; ++l) {
        P.K[l] = pl.K[l]; P.col_of[l] = pl.col_of[l]; P.c2_of[l] = pl.c2_of[l];
        P.masked[l] = (mode == 1 && l > 0 && l < levels - 1) ? 1 : 0;
        P.mask_block[l] = needs[l];
        for (int m = 0; m < l; ++m) P.G[m][l] = w.g[m][l];
    }
    for (int p = 0; p < pl.phases; ++p) P.width[p] = pl.width[p];
    for (int p = 0; p <= pl.phases; ++p) P.first_level[p] = pl.first_level[p];
    P.c2 = w.c2; P.ids = ids; P.flag_list = w.list; P.flag_count = w.count;
    {
        const char* e = getenv("RQK_ENC_FLAG_TOL");
        // the budget's constant was measured at dim = 512 (192 accumulating MMAs per dot product): scale it with the
        // length of the accumulation for longer rows
        P.flag_tol = e ? (float)atof(e) : 4.0f * (dim > 512 ? (float)dim / 512.0f : 1.0f);
    }
    const size_t stage_bytes = 2 * (size_t)TC_BM * TC_BK * 4 + 2 * (size_t)pl.wmax * TC_BK * 4;
    const size_t tail = 8 * 8 * 3 + 2 * 8 * 3 + 2 * 8 * TC_MAX_RAW + 16 + 2 * TC_BM * 4 + EF_MAX_KSUM * 4 + 64;
    const size_t raw_bytes = (size_t)TC_BM * TC_BK * 4;
    const long long room = 225 * 1024 - (long long)tail - 1024;
    int stages = 2;
    if (room < (long long)(stages * stage_bytes + raw_bytes))
        return fail(RQK_ERR_UNSUPPORTED, "rqk_encode_fused: phase width %s%lld does not fit two pipeline stages", "", pl.wmax);
    int raw = (int)((room - (long long)(stages * stage_bytes)) / (long long)raw_bytes);
    if (raw > TC_MAX_RAW) {
        if (room >= (long long)(3 * stage_bytes + 4 * raw_bytes)) {
            stages = 3;
            raw = (int)((room - (long long)(stages * stage_bytes)) / (long long)raw_bytes);
        }
        if (raw > TC_MAX_RAW) raw = TC_MAX_RAW;
    }
    P.stages = stages;
    P.raw = raw;
    const size_t smem = (size_t)stages * stage_bytes + (size_t)raw * raw_bytes + tail + 1024;

    CUtensorMap mx, mhi, mlo;
    int rc;
    if ((rc = make_map_2d(&mx, x, n, dim, TC_BM, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))) return rc;
    if ((rc = make_map_2d(&mhi, w.chi, (long long)pl.phases * pl.wmax, dim, pl.wmax, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))) return rc;
    if ((rc = make_map_2d(&mlo, w.clo, (long long)pl.phases * pl.wmax, dim, pl.wmax, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))) return rc;
    RQK_CUDA_OK(cudaFuncSetAttribute(encode_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    RQK_CUDA_OK(cudaGetDevice(&dev));
    RQK_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long ntiles = ceil_div<long long>(n, TC_BM);
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    encode_fused_kernel<<<grid, TC_THREADS, smem, stream>>>(mx, mhi, mlo, P);
    RQK_LAUNCH_OK();

    EncFixParams F;
    memset(&F, 0, sizeof(F));
    F.n = n; F.dim = dim; F.levels = levels;
    for (int l = 0; l < levels; ++l) {
        F.K[l] = pl.K[l]; F.c2_of[l] = pl.c2_of[l]; F.masked[l] = P.masked[l]; F.mask_block[l] = needs[l];
        F.C[l] = (const float*)centers[l];
    }
    F.c2 = w.c2; F.ids = ids; F.flag_list = w.list; F.flag_count = w.count;
    const size_t fsm = ((size_t)dim + 256 + 8) * 4;
    encode_fixup_kernel<<<sms * 4, EFX_THREADS, fsm, stream>>>(x, F);
    RQK_LAUNCH_OK();
    return 0;
}

}  // extern "C"
