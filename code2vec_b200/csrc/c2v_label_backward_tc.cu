// c2v_label_backward_tc.cu -- backward of the label head  outputs = cv . W_out^T + b  (model.py:83 under loss.backward(),
// main.py:174) on the tensor cores.  G = d loss / d outputs [B, C] fp32 (row-major, any C):
//   LB1  dW_out[c, h] = sum_b G[b, c] cv[b, h]      M = c, N = h, K = b : both operands MN-major (the row index is K)
//        d_b[c]       = sum_b G[b, c]               (column sums, folded into LB1's producers)
//   LB2  d_cv[b, h]   = sum_c G[b, c] W_out[c, h]   M = b, N = h, K = c : A K-major (G rows), B MN-major = the forward's
//        cached W_out image ([128 c x 64 h] fp16 hi/lo tiles) streamed as it is with cp.async.bulk
// Same fp32-accurate scheme as everywhere: operands split into fp16 hi + lo, three kind::f16 MMAs per k-step
// (hi.hi + lo.hi + hi.lo), fp32 accumulation in TMEM; G is multiplied by the power of two that lifts max |G| just
// below 2^14 before the split (mean-NLL gradients are ~1e-6) and the result by its exact inverse.
// The shared-memory image of a G tile is identical for both kernels ([128 rows(b) x 64 cols(c)] SWIZZLE_128B panels);
// only the descriptors differ (MN-major in LB1, K-major in LB2).
// Replaces three CUDA-core launches (two split-K sgemm_kernel + colsum_kernel: 0.32 ms at cfg2, > 2 ms at C = 195,299).
#include <cuda_fp16.h>

#include <cstring>

#include "c2v_tc_ptx.cuh"

namespace c2v {

namespace lbt {
constexpr int ROWS = 128;
constexpr int PANEL = ROWS * 64 * 2;                  // 16 KB: [128 rows x 64 cols] fp16
constexpr int N_PROD_WARPS = 16, PROD_WARP0 = 4, MMA_WARP = 20, BULK_WARP = 21;
constexpr int THREADS = 22 * 32;
constexpr int ROWS_PER_PW = ROWS / N_PROD_WARPS;      // 8
}  // namespace lbt

__device__ __forceinline__ uint64_t lbt_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {   // MN-major SWIZZLE_128B
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float lbt_scale_from_absmax(unsigned bits) {
    const float mx = __uint_as_float(bits);
    if (!(mx > 0.0f && mx < 3.0e38f)) return 1.0f;
    int e;
    frexpf(mx, &e);
    int k = 14 - e;
    k = k > 100 ? 100 : (k < -100 ? -100 : k);
    return ldexpf(1.0f, k);
}
// 4 consecutive floats of a row that may be only 4-byte aligned (C odd) and may run off the end of the row
__device__ __forceinline__ float4 lbt_load4(const float *p, long long remaining, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (remaining >= 4 && vec) return ldg_nc_v4(reinterpret_cast<const float4 *>(p));
    if (remaining > 0) v.x = __ldg(p);
    if (remaining > 1) v.y = __ldg(p + 1);
    if (remaining > 2) v.z = __ldg(p + 2);
    if (remaining > 3) v.w = __ldg(p + 3);
    return v;
}
__device__ __forceinline__ void lbt_split_store(uint32_t hi, uint32_t lo, uint32_t off, float4 v) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
    sts_v2(hi + off, pack_h2(h01), pack_h2(h23));
    sts_v2(lo + off, pack_h2(l01), pack_h2(l23));
}

__global__ void lbt_absmax_kernel(const float *__restrict__ x, long long n, unsigned *__restrict__ out_bits)
{
    float m = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(out_bits, __float_as_uint(m));
}

// ------------------------------------------------------------------------------------------------------------------
// LB1: dW_out [C, H] and d_b [C].  One persistent CTA walks label tiles (128 labels); per label tile all batch tiles
// (128 bags) stream through: A = G[bags, labels] as two 64-label panels, B = cv[bags, 64 h] panels; the [128 c x HP h]
// accumulator (two TMEM stages) is complete after the last batch tile and is written out by warps 0-3.
// smem: 2 A stages {hi p0, hi p1, lo p0, lo p1} (128 KB) + 2 B slots {hi, lo} (64 KB) + column-sum scratch + barriers.
// ------------------------------------------------------------------------------------------------------------------
namespace lb1 {
constexpr int A_STAGE = 4 * lbt::PANEL, B_SLOT = 2 * lbt::PANEL;
constexpr int SMEM_A_OFF = 0, SMEM_B_OFF = 2 * A_STAGE, SMEM_CS_OFF = SMEM_B_OFF + 2 * B_SLOT;   // column sums [16][128] fp32
constexpr int SMEM_BAR_OFF = SMEM_CS_OFF + lbt::N_PROD_WARPS * 128 * 4;
constexpr int SMEM_BYTES = SMEM_BAR_OFF + 128 + 1024;
// kind::f16, D = f32, A and B MN-major, N = 64, M = 128
constexpr uint32_t IDESC = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}  // namespace lb1

// CV_IMG: the B operand (cv) is not converted here at all: `cv_img` is the fp16 hi/lo image of cv that the label GEMM of the
// same step left in the label workspace ([128 bags x 64 h] K-major SWIZZLE_128B tiles = the very bytes an MN-major B tile
// of this GEMM needs), streamed by warp 21 with one 32 KB cp.async.bulk per (batch tile, k-block); the producers then only
// handle G and issue a step's whole G tile (8 pieces per lane) before they wait for the stage.  Per (label tile, batch tile)
// step the producers' chain was four dependent load -> convert -> store rounds (~5-7 us: 623 us for the 1,526 label tiles
// of top11); now one.
template <bool CV_IMG>
__global__ void __launch_bounds__(lbt::THREADS, 1)
label_dw_tc_kernel(const float *__restrict__ G, const float *__restrict__ cv, int B, long long C, int H, int nkb,
                   const unsigned *__restrict__ g_absmax, float *__restrict__ dW, float *__restrict__ d_bias,
                   const uint8_t *__restrict__ cv_img)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char *smem = smem_raw + (base - raw);
    float *s_cs = reinterpret_cast<float *>(smem + lb1::SMEM_CS_OFF);
    const uint32_t bars = base + lb1::SMEM_BAR_OFF;
    // a_full[2] @0, a_empty[2] @16, b_full[2] @32, b_empty[2] @48, acc_full[2] @64, acc_empty[2] @80, tmem ptr @96
    const uint32_t bar_afull = bars, bar_aempty = bars + 16, bar_bfull = bars + 32, bar_bempty = bars + 48,
                   bar_accfull = bars + 64, bar_accempty = bars + 80;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + lb1::SMEM_BAR_OFF + 96);
    __shared__ long long s_status[2];
    long long *status = s_status;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_bt = (B + 127) / 128;
    const long long n_ct = (C + 127) / 128;
    const int my_ct = (int)((n_ct - (long long)blockIdx.x + (long long)gridDim.x - 1) / (long long)gridDim.x);
    const int HP = nkb * 64;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_afull + 8 * s, 2 * lbt::N_PROD_WARPS);
            mbar_init(bar_aempty + 8 * s, 1);
            mbar_init(bar_bfull + 8 * s, CV_IMG ? 1 : lbt::N_PROD_WARPS);
            mbar_init(bar_bempty + 8 * s, 1);
            mbar_init(bar_accfull + 8 * s, 1);
            mbar_init(bar_accempty + 8 * s, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == lbt::MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const float g_scale = lbt_scale_from_absmax(*g_absmax);

    if (warp >= lbt::PROD_WARP0 && warp < lbt::MMA_WARP) {
        // =============================== PRODUCERS ===============================
        const int pw = warp - lbt::PROD_WARP0;
        const int sub_row = lane >> 4, q = lane & 15;   // lanes 0-15: row 2j, lanes 16-31: row 2j+1; q = 16-B column piece
        uint32_t st_off[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = pw * lbt::ROWS_PER_PW + 2 * j + sub_row;
            st_off[j] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((q >> 1) ^ (r & 7)) & 7) << 4) + (q & 1) * 8);
        }
        const bool vecG = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0;
        int ita = 0, itb = 0;                          // running A-stage / B-slot counters
        for (int cl = 0; cl < my_ct; ++cl) {
            const long long ct = (long long)blockIdx.x + (long long)cl * gridDim.x;
            const long long c0 = ct * 128;
            float4 cs[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};   // column sums (unscaled)
            for (int bt = 0; bt < n_bt; ++bt, ++ita) {
                const int row0 = bt * 128 + pw * lbt::ROWS_PER_PW;
                const int as = ita & 1;
                float4 g[2][4];
                if (CV_IMG) {                            // the step's whole G tile in flight before the stage wait
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int r = row0 + 2 * j + sub_row;
                            const long long c = c0 + p * 64 + q * 4;
                            g[p][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (r < B) g[p][j] = lbt_load4(G + (size_t)r * C + c, C - c, vecG);
                        }
                }
                mbar_wait(bar_aempty + 8 * as, (((uint32_t)(ita >> 1)) & 1u) ^ 1u, status);
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const uint32_t hi = base + lb1::SMEM_A_OFF + as * lb1::A_STAGE + p * lbt::PANEL, lo = hi + 2 * lbt::PANEL;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = row0 + 2 * j + sub_row;
                        const long long c = c0 + p * 64 + q * 4;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (CV_IMG) v = g[p][j];
                        else if (r < B) v = lbt_load4(G + (size_t)r * C + c, C - c, vecG);
                        cs[p].x += v.x; cs[p].y += v.y; cs[p].z += v.z; cs[p].w += v.w;
                        v.x *= g_scale; v.y *= g_scale; v.z *= g_scale; v.w *= g_scale;
                        lbt_split_store(hi, lo, st_off[j], v);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_afull + 8 * as);
                }
#pragma unroll 1
                for (int kb = 0; kb < (CV_IMG ? 0 : nkb); ++kb, ++itb) {
                    const int bs = itb & 1;
                    float4 buf[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = row0 + 2 * j + sub_row;
                        const int h = kb * 64 + q * 4;
                        buf[j] = (r < B && h < H) ? ldg_nc_v4(reinterpret_cast<const float4 *>(cv + (size_t)r * H + h))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    mbar_wait(bar_bempty + 8 * bs, (((uint32_t)(itb >> 1)) & 1u) ^ 1u, status);
                    const uint32_t hi = base + lb1::SMEM_B_OFF + bs * lb1::B_SLOT;
#pragma unroll
                    for (int j = 0; j < 4; ++j) lbt_split_store(hi, hi + lbt::PANEL, st_off[j], buf[j]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_bfull + 8 * bs);
                }
            }
            // ---- d_b of this label tile: lanes q / q+16 hold the same columns; 16 warps -> smem -> warp 0 of the producers
            if (d_bias) {
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    cs[p].x += __shfl_xor_sync(0xffffffffu, cs[p].x, 16); cs[p].y += __shfl_xor_sync(0xffffffffu, cs[p].y, 16);
                    cs[p].z += __shfl_xor_sync(0xffffffffu, cs[p].z, 16); cs[p].w += __shfl_xor_sync(0xffffffffu, cs[p].w, 16);
                    if (sub_row == 0) *reinterpret_cast<float4 *>(s_cs + pw * 128 + p * 64 + q * 4) = cs[p];
                }
                named_bar_sync(2, lbt::N_PROD_WARPS * 32);
                const int t = (warp - lbt::PROD_WARP0) * 32 + lane;     // 0..511
                if (t < 128) {
                    float s = 0.0f;
#pragma unroll
                    for (int w = 0; w < lbt::N_PROD_WARPS; ++w) s += s_cs[w * 128 + t];
                    if (c0 + t < C) d_bias[c0 + t] = s;
                }
                named_bar_sync(2, lbt::N_PROD_WARPS * 32);                // scratch is rewritten by the next label tile
            }
        }
    } else if (CV_IMG && warp == lbt::BULK_WARP) {
        // =============================== cv IMAGE PRODUCER (cp.async.bulk) ===============================
        if (lane == 0) {
            int itb = 0;
            for (int cl = 0; cl < my_ct; ++cl)
                for (int bt = 0; bt < n_bt; ++bt)
                    for (int kb = 0; kb < nkb; ++kb, ++itb) {
                        const int bs = itb & 1;
                        mbar_wait(bar_bempty + 8 * bs, (((uint32_t)(itb >> 1)) & 1u) ^ 1u, status);
                        mbar_arrive_expect_tx(bar_bfull + 8 * bs, lb1::B_SLOT);
                        bulk_copy_g2s(base + lb1::SMEM_B_OFF + bs * lb1::B_SLOT, cv_img + ((size_t)bt * nkb + kb) * lb1::B_SLOT,
                                      lb1::B_SLOT, bar_bfull + 8 * bs);
                    }
        }
        __syncwarp();
    } else if (warp == lbt::MMA_WARP) {
        // =============================== MMA ISSUER (converged, one elected lane) ===============================
        int ita = 0, itb = 0;
        for (int cl = 0; cl < my_ct; ++cl) {
            const int acs = cl & 1;
            mbar_wait(bar_accempty + 8 * acs, (((uint32_t)(cl >> 1)) & 1u) ^ 1u, status);
            for (int bt = 0; bt < n_bt; ++bt, ++ita) {
                const int as = ita & 1;
                mbar_wait(bar_afull + 8 * as, ((uint32_t)(ita >> 1)) & 1u, status);
#pragma unroll 1
                for (int kb = 0; kb < nkb; ++kb, ++itb) {
                    const int bs = itb & 1;
                    mbar_wait(bar_bfull + 8 * bs, ((uint32_t)(itb >> 1)) & 1u, status);
                    fence_proxy_async_smem();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = base + lb1::SMEM_A_OFF + as * lb1::A_STAGE;
                        const uint32_t sb = base + lb1::SMEM_B_OFF + bs * lb1::B_SLOT;
                        const uint32_t d_tmem = tmem_base + (uint32_t)(acs * 256 + kb * 64);
#pragma unroll
                        for (int k = 0; k < lbt::ROWS / 16; ++k) {
                            const uint64_t a_hi = lbt_desc_mn(sa + k * 2048, lbt::PANEL);
                            const uint64_t a_lo = lbt_desc_mn(sa + 2 * lbt::PANEL + k * 2048, lbt::PANEL);
                            const uint64_t b_hi = lbt_desc_mn(sb + k * 2048, lbt::PANEL);
                            const uint64_t b_lo = lbt_desc_mn(sb + lbt::PANEL + k * 2048, lbt::PANEL);
                            umma_f16(d_tmem, a_hi, b_hi, lb1::IDESC, (bt | k) != 0 ? 1u : 0u);
                            umma_f16(d_tmem, a_lo, b_hi, lb1::IDESC, 1u);
                            umma_f16(d_tmem, a_hi, b_lo, lb1::IDESC, 1u);
                        }
                        umma_commit(bar_bempty + 8 * bs);
                        if (kb == nkb - 1) {
                            umma_commit(bar_aempty + 8 * as);
                            if (bt == n_bt - 1) umma_commit(bar_accfull + 8 * acs);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 4) {
        // =============================== DRAIN: accumulator -> dW_out rows ===============================
        const float inv = 1.0f / g_scale;
        for (int cl = 0; cl < my_ct; ++cl) {
            const int acs = cl & 1;
            const long long c = ((long long)blockIdx.x + (long long)cl * gridDim.x) * 128 + warp * 32 + lane;
            mbar_wait(bar_accfull + 8 * acs, ((uint32_t)(cl >> 1)) & 1u, status);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < HP / 32; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acs * 256 + ch * 32), v);
                tmem_ld_wait();
                if (c < C) {
                    float *dst = dW + (size_t)c * H + ch * 32;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (ch * 32 + j < H)
                            *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j] * inv, v[j + 1] * inv, v[j + 2] * inv, v[j + 3] * inv);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty + 8 * acs);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == lbt::MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// LB2: d_cv [B, H] += G[bags, label range] . W_out[label range, :]   (split over the label dimension, fp32 atomics).
// Work item = (batch tile, label split); k-block = 64 labels: A = one K-major G panel {hi, lo} (32 KB, 3 stages),
// B = rows [64 half] of the forward's W_out image tiles, nkb x {hi 8 KB, lo 8 KB} (3 stages), UMMA N = HP.
// ------------------------------------------------------------------------------------------------------------------
namespace lb2 {
constexpr int MAX_STAGES = 3;
constexpr int A_STAGE = 2 * lbt::PANEL;               // hi | lo, [128 bags x 64 labels]
constexpr int B_HALF = 64 * 128;                      // 8 KB: 64 label rows x 64 h fp16
constexpr int SMEM_A_OFF = 0;
}  // namespace lb2
// the B stage is nkb x {hi, lo} x 8 KB: 3 stages of (32 + 32) KB at encode_size <= 128, 2 stages of (32 + 64) KB at 256
static inline int lb2_stages(int nkb) { return nkb <= 2 ? 3 : 2; }
static inline int lb2_smem_bytes(int nkb) { return lb2_stages(nkb) * (lb2::A_STAGE + nkb * 2 * lb2::B_HALF) + 128 + 1024; }

__global__ void __launch_bounds__(lbt::THREADS, 1)
label_dcv_tc_kernel(const float *__restrict__ G, int B, long long C, int H, int nkb, const uint8_t *__restrict__ w_img,
                    const float *__restrict__ w_hdr, const unsigned *__restrict__ g_absmax, int n_split,
                    long long kblocks_per_split, float *__restrict__ d_cv, const int S)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char *smem = smem_raw + (base - raw);
    const int b_stage = nkb * 2 * lb2::B_HALF;
    const uint32_t smem_b = base + S * lb2::A_STAGE;
    const uint32_t bars = smem_b + S * b_stage;
    // a_full[3] @0, a_empty[3] @24, b_full[3] @48, b_empty[3] @72, acc_full @96, acc_empty @104, tmem ptr @112
    const uint32_t bar_afull = bars, bar_aempty = bars + 24, bar_bfull = bars + 48, bar_bempty = bars + 72,
                   bar_accfull = bars + 96, bar_accempty = bars + 104;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + (bars - base) + 112);
    __shared__ long long s_status[2];
    long long *status = s_status;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_bt = (B + 127) / 128;
    const int n_items = n_bt * n_split;
    const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const long long n_kblocks = (C + 63) / 64;
    const int HP = nkb * 64;
    // kind::f16, D = f32, A K-major, B MN-major, N = HP, M = 128
    const uint32_t idesc = (1u << 4) | (1u << 16) | ((uint32_t)(HP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    if (tid == 0) {
        for (int s = 0; s < lb2::MAX_STAGES; ++s) {
            mbar_init(bar_afull + 8 * s, lbt::N_PROD_WARPS);
            mbar_init(bar_aempty + 8 * s, 1);
            mbar_init(bar_bfull + 8 * s, 1);
            mbar_init(bar_bempty + 8 * s, 1);
        }
        mbar_init(bar_accfull, 1);
        mbar_init(bar_accempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == lbt::MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const float g_scale = lbt_scale_from_absmax(*g_absmax);

    // item -> (batch tile, first k-block, number of k-blocks)
    auto item_range = [&](int il, int &bt, long long &kb0, long long &nk) {
        const int item = (int)blockIdx.x + il * (int)gridDim.x;
        bt = item / n_split;
        const int sp = item - bt * n_split;
        kb0 = (long long)sp * kblocks_per_split;
        nk = n_kblocks - kb0; if (nk > kblocks_per_split) nk = kblocks_per_split; if (nk < 0) nk = 0;
    };

    if (warp >= lbt::PROD_WARP0 && warp < lbt::MMA_WARP) {
        // =============================== G PRODUCERS (K-major panels) ===============================
        const int pw = warp - lbt::PROD_WARP0;
        const int sub_row = lane >> 4, q = lane & 15;
        uint32_t st_off[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = pw * lbt::ROWS_PER_PW + 2 * j + sub_row;
            st_off[j] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((q >> 1) ^ (r & 7)) & 7) << 4) + (q & 1) * 8);
        }
        const bool vecG = (C & 3) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0;
        int it = 0;
        for (int il = 0; il < my_items; ++il) {
            int bt; long long kb0, nk;
            item_range(il, bt, kb0, nk);
            const int row0 = bt * 128 + pw * lbt::ROWS_PER_PW;
            for (long long kk = 0; kk < nk; ++kk, ++it) {
                const long long c = (kb0 + kk) * 64 + q * 4;
                float4 buf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int r = row0 + 2 * j + sub_row;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < B) v = lbt_load4(G + (size_t)r * C + c, C - c, vecG);
                    v.x *= g_scale; v.y *= g_scale; v.z *= g_scale; v.w *= g_scale;
                    buf[j] = v;
                }
                const int st = it % S;
                mbar_wait(bar_aempty + 8 * st, (((uint32_t)(it / S)) & 1u) ^ 1u, status);
                const uint32_t hi = base + lb2::SMEM_A_OFF + st * lb2::A_STAGE;
#pragma unroll
                for (int j = 0; j < 4; ++j) lbt_split_store(hi, hi + lbt::PANEL, st_off[j], buf[j]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * st);
            }
        }
    } else if (warp == lbt::BULK_WARP) {
        // =============================== W_out IMAGE PRODUCER (cp.async.bulk) ===============================
        if (lane == 0) {
            int it = 0;
            for (int il = 0; il < my_items; ++il) {
                int bt; long long kb0, nk;
                item_range(il, bt, kb0, nk);
                for (long long kk = 0; kk < nk; ++kk, ++it) {
                    const long long kblk = kb0 + kk;                 // 64-label block
                    const long long nt = kblk >> 1;                  // 128-label tile of the image
                    const int half = (int)(kblk & 1);
                    const int st = it % S;
                    mbar_wait(bar_bempty + 8 * st, (((uint32_t)(it / S)) & 1u) ^ 1u, status);
                    mbar_arrive_expect_tx(bar_bfull + 8 * st, (uint32_t)b_stage);
                    for (int kb = 0; kb < nkb; ++kb) {
                        const uint8_t *src = w_img + ((size_t)nt * nkb + kb) * (size_t)(2 * lbt::PANEL) + half * lb2::B_HALF;
                        const uint32_t dst = smem_b + st * b_stage + kb * 2 * lb2::B_HALF;
                        bulk_copy_g2s(dst, src, lb2::B_HALF, bar_bfull + 8 * st);                                // hi
                        bulk_copy_g2s(dst + lb2::B_HALF, src + lbt::PANEL, lb2::B_HALF, bar_bfull + 8 * st);     // lo
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == lbt::MMA_WARP) {
        // =============================== MMA ISSUER ===============================
        int it = 0;
        for (int il = 0; il < my_items; ++il) {
            int bt; long long kb0, nk;
            item_range(il, bt, kb0, nk);
            mbar_wait(bar_accempty, ((uint32_t)il & 1u) ^ 1u, status);
            for (long long kk = 0; kk < nk; ++kk, ++it) {
                const int st = it % S;
                const uint32_t ph = ((uint32_t)(it / S)) & 1u;
                mbar_wait(bar_afull + 8 * st, ph, status);
                mbar_wait(bar_bfull + 8 * st, ph, status);
                fence_proxy_async_smem();
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = base + lb2::SMEM_A_OFF + st * lb2::A_STAGE;
                    const uint32_t sb = smem_b + st * b_stage;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                   // 64 labels = 4 k-steps of 16
                        const uint64_t a_hi = umma_desc(sa + k * 32), a_lo = umma_desc(sa + lbt::PANEL + k * 32);
                        // B: MN-major, 16 label rows x 128 B per k-step, 64-column groups 2 * B_HALF apart
                        const uint64_t b_hi = lbt_desc_mn(sb + k * 2048, 2 * lb2::B_HALF);
                        const uint64_t b_lo = lbt_desc_mn(sb + lb2::B_HALF + k * 2048, 2 * lb2::B_HALF);
                        umma_f16(tmem_base, a_hi, b_hi, idesc, (kk | k) != 0 ? 1u : 0u);
                        umma_f16(tmem_base, a_lo, b_hi, idesc, 1u);
                        umma_f16(tmem_base, a_hi, b_lo, idesc, 1u);
                    }
                    umma_commit(bar_aempty + 8 * st);
                    umma_commit(bar_bempty + 8 * st);
                    if (kk == nk - 1) umma_commit(bar_accfull);
                }
                __syncwarp();
            }
            if (nk == 0 && elect_one()) {                           // empty item: nothing to add, release the epilogue
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_accfull) : "memory");
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        // =============================== EPILOGUE: accumulator -> atomics into d_cv ===============================
        const float inv = w_hdr[0] / g_scale;
        for (int il = 0; il < my_items; ++il) {
            int bt; long long kb0, nk;
            item_range(il, bt, kb0, nk);
            const int b = bt * 128 + warp * 32 + lane;
            mbar_wait(bar_accfull, (uint32_t)il & 1u, status);
            tc_fence_after();
            if (nk > 0) {
#pragma unroll 1
                for (int ch = 0; ch < HP / 32; ++ch) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ch * 32), v);
                    tmem_ld_wait();
                    if (b < B) {
                        float *dst = d_cv + (size_t)b * H + ch * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            if (ch * 32 + j < H)
                                red_add_v4(dst + j, make_float4(v[j] * inv, v[j + 1] * inv, v[j + 2] * inv, v[j + 3] * inv));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == lbt::MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

bool label_backward_tc_ok(const c2v_dims *d)
{
    return d->encode >= 4 && d->encode <= 256 && (d->encode & 3) == 0;
}

// scratch: 256 B (word 0: bits of max |G|).  w_img / w_hdr: the label workspace's cached W_out image and header (valid).
int launch_label_backward_tc(const c2v_dims *d, const float *cv, const float *G, int B, const uint8_t *w_img,
                             const float *w_hdr, float *d_cv, float *d_w, float *d_b, unsigned *scratch, cudaStream_t st,
                             bool absmax_ready, const uint8_t *cv_img)
{
    const int H = d->encode, nkb = (H + 63) / 64;
    const long long C = d->label_count;
    int dev = 0, sms = 0;
    C2V_CUDA_OK(cudaGetDevice(&dev));
    C2V_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (!absmax_ready) {                       // (ready: c2v_label_dlogits left max |G| in the word while it wrote G)
        C2V_CUDA_OK(cudaMemsetAsync(scratch, 0, 4, st));
        lbt_absmax_kernel<<<sms * 4, 256, 0, st>>>(G, (long long)B * C, scratch);
        C2V_LAUNCH_OK("lbt_absmax_kernel");
    }
    if (d_w || d_b) {
        if (!d_w) { set_error("label backward (tensor cores): d_output_bias needs d_output_weight"); return C2V_EINVAL; }
        auto kern = cv_img ? label_dw_tc_kernel<true> : label_dw_tc_kernel<false>;
        C2V_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lb1::SMEM_BYTES));
        const long long n_ct = (C + 127) / 128;
        const int grid = (int)(n_ct < sms ? n_ct : sms);
        kern<<<grid, lbt::THREADS, lb1::SMEM_BYTES, st>>>(G, cv, B, C, H, nkb, scratch, d_w, d_b, cv_img);
        C2V_LAUNCH_OK("label_dw_tc_kernel");
    }
    if (d_cv) {
        C2V_CUDA_OK(cudaMemsetAsync(d_cv, 0, (size_t)B * H * sizeof(float), st));
        const int smem = lb2_smem_bytes(nkb);
        C2V_CUDA_OK(cudaFuncSetAttribute(label_dcv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int n_bt = (B + 127) / 128;
        const long long n_kblocks = (C + 63) / 64;
        long long n_split = (2LL * sms + n_bt - 1) / n_bt;            // ~2 items per SM
        if (n_split > n_kblocks) n_split = n_kblocks;
        if (n_split < 1) n_split = 1;
        const long long per = (n_kblocks + n_split - 1) / n_split;
        n_split = (n_kblocks + per - 1) / per;
        const long long n_items = (long long)n_bt * n_split;
        const int grid = (int)(n_items < sms ? n_items : sms);
        label_dcv_tc_kernel<<<grid, lbt::THREADS, smem, st>>>(G, B, C, H, nkb, w_img, w_hdr, scratch, (int)n_split, per, d_cv,
                                                              lb2_stages(nkb));
        C2V_LAUNCH_OK("label_dcv_tc_kernel");
    }
    return C2V_OK;
}

}  // namespace c2v
