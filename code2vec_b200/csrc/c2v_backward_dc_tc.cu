// c2v_backward_dc_tc.cu -- K3c: dC = dX . W on the tensor cores + scatter into the embedding gradients
// (terminal_embed = path_embed = E <= 256, encode = H <= 256, multiples of 4; zero-padded to 128 inside the panels).
//
// The gradient of the gathered context vectors (autograd of model.py:48-54 under loss.backward(), main.py:174):
//   dC[r, d] = sum_h dX[r, h] * W[h, d],  then  dE_t[starts_r] += dC[r, 0:128], dE_p[paths_r] += dC[r, 128:256],
//   dE_t[ends_r] += dC[r, 256:384]   (nn.Embedding's dense backward; PAD row 0 is a learned row and gets its share).
// M = 128 context rows per tile (TMEM lanes), N = 384 = three 128-column accumulators (one per sub-vector), K = 128.
// A = dX, K-major: the same [128 rows x 64 h] fp16 SWIZZLE_128B hi/lo panels the dW kernel builds (c2v_backward_dw_tc.cu),
// read here with K-major descriptors.  B = W^T as a K-major image [128 d x 64 h] per (sub-vector, k-block), built once
// per call by split_wt_kernel and streamed from L2 with 32 KB cp.async.bulk copies.  3-pass fp16 hi/lo split, fp32
// accumulation in TMEM; dX is pre-scaled by the power of two from max |dx|, W by the one from max |W|.
//
// Warps: 0-3 scatter epilogue (thread = context row: tcgen05.ld -> 128-bit atomics into the three embedding rows) |
// 4-19 dX producers | 20 MMA issuer | 21 W^T producer.  The three accumulators are released one by one, so the scatter
// of sub-vector sv overlaps the MMAs of sv+1 (and of the next tile).
// Sizes above 128: blockIdx.y = db selects the 128-wide window of d inside every sub-vector (its own grid of CTAs over the
// row tiles); the contraction over h runs as n_hb blocks of 128 through the same operand stages, accumulating in tensor
// memory, and the accumulator is handed to the scatter warps after the last block -- so every dC element is still added
// to the embedding gradient exactly once.
#include <cuda_fp16.h>

#include "c2v_tc_ptx.cuh"

namespace c2v {

namespace dct {
constexpr int ROWS = 128, H = 128, E = 128, D = 3 * E;
constexpr int PANEL = ROWS * 64 * 2;                  // 16 KB
constexpr int A_STAGE = 4 * PANEL;                    // hi k0 | hi k1 | lo k0 | lo k1   (k-block = 64 h)
constexpr int B_SLOT = 2 * PANEL;                     // hi | lo of one (sub-vector, k-block) tile of W^T
constexpr int N_PROD_WARPS = 16, PROD_WARP0 = 4, MMA_WARP = 20, W_WARP = 21;
constexpr int THREADS = 22 * 32;
constexpr int ROWS_PER_PW = ROWS / N_PROD_WARPS;      // 8
constexpr int NB = 6;                                 // W^T tiles per row tile: (sv, kb), kb minor
constexpr int SMEM_A_OFF = 0, SMEM_B_OFF = 2 * A_STAGE, SMEM_BAR_OFF = SMEM_B_OFF + 2 * B_SLOT;
constexpr int STG_LD = 36;                            // padded fp32 row of a scatter warp's [32 rows x 32 columns] staging tile
constexpr int SMEM_STG_OFF = SMEM_BAR_OFF + 128;
constexpr int SMEM_BYTES = SMEM_STG_OFF + 4 * 32 * STG_LD * 4 + 1024;
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);   // K-major A and B
constexpr int IMG_BYTES = NB * B_SLOT;                // 192 KB
}  // namespace dct

// W [H][D = 3E] fp32 -> per (db, hb) window pair 6 tiles (sv * 2 + kb) of {hi, lo} [128 d x 64 h] fp16, K-major
// SWIZZLE_128B (d = 128 db + row, h = 128 hb + 64 kb + column; zeros beyond E / H), scaled by the power of two that lifts
// max |W| (bits in *absmax_bits, found by wt_absmax_kernel) just below 2^14.  Image order: [db][hb][sv][kb].
__global__ void wt_absmax_kernel(const float *__restrict__ W, int n, unsigned *__restrict__ absmax_bits)
{
    float m = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(W[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(absmax_bits, __float_as_uint(m));
}
__global__ void __launch_bounds__(256)
split_wt_kernel(const float *__restrict__ W, int H, int E, int n_hb, int n_db, const unsigned *__restrict__ absmax_bits,
                uint8_t *__restrict__ img, float *__restrict__ hdr)
{
    const float mx = __uint_as_float(*absmax_bits);
    float scale = 1.0f;
    if (mx > 0.0f && mx < 3.0e38f) {
        int e;
        frexpf(mx, &e);
        int k = 14 - e;
        k = k > 60 ? 60 : (k < -60 ? -60 : k);
        scale = ldexpf(1.0f, k);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { hdr[0] = 1.0f / scale; hdr[1] = scale; }
    // one thread = 4 consecutive d of one h of one window pair's PADDED [128 h][3 x 128 d] matrix (zeros beyond H / E);
    // coalesced 16-B reads of W's rows ([H][3E] row-major)
    constexpr int PER_WIN = dct::H * dct::D / 4;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_db * n_hb * PER_WIN; g += gridDim.x * blockDim.x) {
        const int win = g / PER_WIN, gw = g % PER_WIN;      // win = db * n_hb + hb
        const int db = win / n_hb, hb = win % n_hb;
        const int hl = gw / (dct::D / 4), dp = (gw % (dct::D / 4)) * 4;
        const int sv = dp / dct::E, dl = dp % dct::E, kb = hl / 64, kk = hl % 64;
        const int h = hb * dct::H + hl, d = db * dct::E + dl;
        float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (h < H && d < E) w4 = *reinterpret_cast<const float4 *>(W + (size_t)h * (3 * E) + sv * E + d);
        const float wv[4] = {w4.x * scale, w4.y * scale, w4.z * scale, w4.w * scale};
        uint8_t *base = img + ((size_t)win * dct::NB + sv * 2 + kb) * dct::B_SLOT;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const __half hi = __float2half_rn(wv[q]);
            const __half lo = __float2half_rn(wv[q] - __half2float(hi));
            const uint32_t off = sw128_offset(dl + q, kk);
            *reinterpret_cast<__half *>(base + off) = hi;
            *reinterpret_cast<__half *>(base + dct::PANEL + off) = lo;
        }
    }
}

__global__ void __launch_bounds__(dct::THREADS, 1)
backward_dc_tc_kernel(const EncodeArgs a, const float *__restrict__ dx, const unsigned *__restrict__ dx_absmax,
                      const uint8_t *__restrict__ wt_img, const float *__restrict__ wt_hdr,
                      float *__restrict__ g_emb_t, float *__restrict__ g_emb_p, const int sv_mask, const int n_hb)
{
    const int db = (int)blockIdx.y;                     // this CTA's 128-wide window of d inside every sub-vector
    // sv_mask: which sub-vectors (bit 0 start, 1 path, 2 end) this launch handles.  The training step runs the path
    // sub-vector first (mask 2): the path table's gradient is then complete and its data-parallel reduction can overlap
    // the start / end launch (mask 5) -- see ShardedFlatAdam.early_step.
    const int last_sv = (sv_mask & 4) ? 2 : ((sv_mask & 2) ? 1 : 0);
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char *smem = smem_raw + (base - raw);
    const uint32_t bars = base + dct::SMEM_BAR_OFF;
    // a_full[2] @0, a_empty[2] @16, b_full[2] @32, b_empty[2] @48, t_full[3] @64, t_empty[3] @88, tmem ptr @112
    const uint32_t bar_afull = bars, bar_aempty = bars + 16, bar_bfull = bars + 32, bar_bempty = bars + 48,
                   bar_tfull = bars + 64, bar_tempty = bars + 88;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + dct::SMEM_BAR_OFF + 112);
    __shared__ long long s_status[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    long long *status = s_status;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_afull + 8 * s, 2 * dct::N_PROD_WARPS);
            mbar_init(bar_aempty + 8 * s, 1);
            mbar_init(bar_bfull + 8 * s, 1);
            mbar_init(bar_bempty + 8 * s, 1);
        }
        for (int s = 0; s < 3; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == dct::MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    float dx_scale = 1.0f;                              // power of two that lifts max |dx| just below 2^14 (see K3b)
    {
        const float mx = __uint_as_float(*dx_absmax);
        if (mx > 0.0f && mx < 3.0e38f) {
            int e;
            frexpf(mx, &e);
            int k = 14 - e;
            k = k > 100 ? 100 : (k < -100 ? -100 : k);
            dx_scale = ldexpf(1.0f, k);
        }
    }

    if (warp >= dct::PROD_WARP0 && warp < dct::MMA_WARP) {
        // =============================== dX PRODUCERS ===============================
        const int pw = warp - dct::PROD_WARP0;
        const int sub_row = lane >> 4, q = lane & 15;
        uint32_t st_off[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = pw * dct::ROWS_PER_PW + 2 * j + sub_row;
            st_off[j] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((q >> 1) ^ (r & 7)) & 7) << 4) + (q & 1) * 8);
        }
        const float4 *dx4 = reinterpret_cast<const float4 *>(dx);
        const int H4 = a.H / 4;
        for (int vt = 0; vt < my_tiles * n_hb; ++vt) {                 // vt = (row tile, h block): one operand stage each
            const int tl = vt / n_hb, h04 = (vt % n_hb) * 32;          // h block start in 16-byte pieces
            const long long row0 = ((long long)blockIdx.x + (long long)tl * gridDim.x) * dct::ROWS + pw * dct::ROWS_PER_PW;
            const int as = vt & 1;
            float4 buf[2][4];
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long r = row0 + 2 * j + sub_row;
                    float4 v = (r < a.N && h04 + p * 16 + q < H4) ? ldg_nc_v4(dx4 + (size_t)r * H4 + h04 + p * 16 + q)
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                    v.x *= dx_scale; v.y *= dx_scale; v.z *= dx_scale; v.w *= dx_scale;
                    buf[p][j] = v;
                }
            mbar_wait(bar_aempty + 8 * as, (((uint32_t)(vt >> 1)) & 1u) ^ 1u, status);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const uint32_t hi = base + dct::SMEM_A_OFF + as * dct::A_STAGE + p * dct::PANEL, lo = hi + 2 * dct::PANEL;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 v = buf[p][j];
                    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y);
                    const __half2 l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
                    sts_v2(hi + st_off[j], pack_h2(h01), pack_h2(h23));
                    sts_v2(lo + st_off[j], pack_h2(l01), pack_h2(l23));
                }
                fence_proxy_async_smem();             // writer side: generic-proxy stores -> visible to the tensor core's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * as);
            }
        }
    } else if (warp == dct::W_WARP) {
        // =============================== W^T PRODUCER ===============================
        if (lane == 0) {
            int it = 0;                                  // same (row tile, h block, sub-vector, k-block) order as the MMA warp
            for (int vt = 0; vt < my_tiles * n_hb; ++vt) {
                const uint8_t *win = wt_img + (size_t)(db * n_hb + vt % n_hb) * dct::IMG_BYTES;
                for (int kb6 = 0; kb6 < dct::NB; ++kb6) {
                    if (!((sv_mask >> (kb6 >> 1)) & 1)) continue;
                    const int bs = it & 1;
                    mbar_wait(bar_bempty + 8 * bs, (((uint32_t)(it >> 1)) & 1u) ^ 1u, status);
                    mbar_arrive_expect_tx(bar_bfull + 8 * bs, dct::B_SLOT);
                    bulk_copy_g2s(base + dct::SMEM_B_OFF + bs * dct::B_SLOT, win + (size_t)kb6 * dct::B_SLOT, dct::B_SLOT, bar_bfull + 8 * bs);
                    ++it;
                }
            }
        }
        __syncwarp();
    } else if (warp == dct::MMA_WARP) {
        // =============================== MMA ISSUER (converged, one elected lane) ===============================
        int it = 0;
        for (int vt = 0; vt < my_tiles * n_hb; ++vt) {
            const int tl = vt / n_hb, hb = vt % n_hb;
            const int as = vt & 1;
            mbar_wait(bar_afull + 8 * as, ((uint32_t)(vt >> 1)) & 1u, status);
#pragma unroll 1
            for (int sv = 0; sv < 3; ++sv) {
                if (!((sv_mask >> sv) & 1)) continue;
                if (hb == 0) mbar_wait(bar_tempty + 8 * sv, ((uint32_t)tl & 1u) ^ 1u, status);   // scatter of the previous tile drained
#pragma unroll 1
                for (int kb = 0; kb < 2; ++kb, ++it) {
                    const int bs = it & 1;
                    mbar_wait(bar_bfull + 8 * bs, ((uint32_t)(it >> 1)) & 1u, status);
                    fence_proxy_async_smem();         // producers' generic-proxy stores of the dX panels -> async proxy
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sa = base + dct::SMEM_A_OFF + as * dct::A_STAGE + kb * dct::PANEL;
                        const uint32_t sb = base + dct::SMEM_B_OFF + bs * dct::B_SLOT;
                        const uint32_t d_tmem = tmem_base + (uint32_t)(sv * 128);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t a_hi = umma_desc(sa + k * 32), a_lo = umma_desc(sa + 2 * dct::PANEL + k * 32);
                            const uint64_t b_hi = umma_desc(sb + k * 32), b_lo = umma_desc(sb + dct::PANEL + k * 32);
                            umma_f16(d_tmem, a_hi, b_hi, dct::IDESC, (hb | kb | k) != 0 ? 1u : 0u);
                            umma_f16(d_tmem, a_lo, b_hi, dct::IDESC, 1u);
                            umma_f16(d_tmem, a_hi, b_lo, dct::IDESC, 1u);
                        }
                        umma_commit(bar_bempty + 8 * bs);
                        if (kb == 1) {
                            if (hb == n_hb - 1) umma_commit(bar_tfull + 8 * sv);          // contraction over all of h complete
                            if (sv == last_sv) umma_commit(bar_aempty + 8 * as);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // =============================== SCATTER EPILOGUE ===============================
        // tcgen05.ld hands every thread 32 columns of ITS row; scattered like that, one warp instruction touches 32
        // different embedding rows with 16 bytes each (32 half-filled sectors).  Each 32 x 32 chunk therefore goes through
        // a padded shared-memory tile and leaves as 8 instructions of 4 rows x 128 contiguous bytes (whole sectors, one
        // cache line per row piece): half as many sector-sized reductions for the L2, an eighth of the lines per request.
        const float inv = wt_hdr[0] / dx_scale;        // exact: both scales are powers of two
        float *stg = reinterpret_cast<float *>(smem + dct::SMEM_STG_OFF) + warp * (32 * dct::STG_LD);
        const int wr = lane >> 3, cp = lane & 7;       // writer role: row 4 * it + wr of the warp's 32, 16-byte piece cp of the chunk
        const int E4 = a.Et / 4;
        for (int tl = 0; tl < my_tiles; ++tl) {
            const long long row = ((long long)blockIdx.x + (long long)tl * gridDim.x) * dct::ROWS + warp * 32 + lane;
            long long is = 0, ip = 0, ie = 0;
            const bool in_range = row < a.N;
            if (in_range) { is = a.starts[row]; ip = a.paths[row]; ie = a.ends[row]; }
            if (is < 0 || is >= a.T) is = 0;
            if (ip < 0 || ip >= a.P) ip = 0;
            if (ie < 0 || ie >= a.T) ie = 0;
#pragma unroll 1
            for (int sv = 0; sv < 3; ++sv) {
                if (!((sv_mask >> sv) & 1)) continue;
                // start of this lane's row in the table, in 16-byte units (backward_dc_tc_ok: fits 32 bits); ~0 = no row
                const long long idx = sv == 0 ? is : (sv == 1 ? ip : ie);
                const uint32_t my_off4 = in_range ? (uint32_t)(idx * E4) : 0xFFFFFFFFu;
                uint32_t ro[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) ro[it] = __shfl_sync(0xffffffffu, my_off4, 4 * it + wr);
                float4 *tab4 = reinterpret_cast<float4 *>(sv == 1 ? g_emb_p : g_emb_t) + db * (dct::E / 4);
                mbar_wait(bar_tfull + 8 * sv, (uint32_t)tl & 1u, status);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sv * 128 + c * 32), v);
                    tmem_ld_wait();
                    if (c == 3) {                       // all 128 columns of this accumulator are in registers / done
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * sv);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4 *>(stg + lane * dct::STG_LD + j) =
                            make_float4(v[j] * inv, v[j + 1] * inv, v[j + 2] * inv, v[j + 3] * inv);
                    __syncwarp();
                    const bool col_ok = db * dct::E + c * 32 + cp * 4 < a.Et;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const float4 w = *reinterpret_cast<const float4 *>(stg + (4 * it + wr) * dct::STG_LD + cp * 4);
                        if (ro[it] != 0xFFFFFFFFu && col_ok && (w.x != 0.0f || w.y != 0.0f || w.z != 0.0f || w.w != 0.0f))   // padded contexts: dx == 0
                            red_add_v4(reinterpret_cast<float *>(tab4 + ro[it] + c * 8 + cp), w);
                    }
                    __syncwarp();                       // the staging tile is rewritten by the next chunk
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == dct::MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

bool backward_dc_tc_ok(const EncodeArgs &a) {
    return a.Et == a.Ep && a.Et <= 2 * dct::E && a.H <= 2 * dct::H && (a.Et & 3) == 0 && (a.H & 3) == 0 &&
           (long long)a.T * (a.Et / 4) < 0xFFFFFFFFll && (long long)a.P * (a.Et / 4) < 0xFFFFFFFFll;
}
size_t backward_dc_tc_workspace_bytes() { return 1024 + 4 * dct::IMG_BYTES; }      // up to 2 x 2 window pairs

// ws: [0, 1024) header {1/scale, scale} | W^T images [db][hb]
int launch_backward_dc_tc(const EncodeArgs &a_in, const float *W, const float *dx, const unsigned *dx_absmax, void *ws,
                          float *g_emb_t, float *g_emb_p, cudaStream_t st, int sv_mask, bool build_image)
{
    EncodeArgs a = a_in;
    a.n_tiles = (int)((a.N + dct::ROWS - 1) / dct::ROWS);
    float *hdr = static_cast<float *>(ws);
    uint8_t *img = static_cast<uint8_t *>(ws) + 1024;
    const int n_hb = (a.H + dct::H - 1) / dct::H, n_db = (a.Et + dct::E - 1) / dct::E;
    unsigned *mxbits = reinterpret_cast<unsigned *>(static_cast<uint8_t *>(ws) + 512);
    if (build_image) {
        C2V_CUDA_OK(cudaMemsetAsync(mxbits, 0, 4, st));
        wt_absmax_kernel<<<48, 256, 0, st>>>(W, a.H * a.D, mxbits);
        C2V_LAUNCH_OK("wt_absmax_kernel");
        split_wt_kernel<<<48 * n_hb * n_db, 256, 0, st>>>(W, a.H, a.Et, n_hb, n_db, mxbits, img, hdr);
        C2V_LAUNCH_OK("split_wt_kernel");
    }
    if ((sv_mask & 7) == 0) return C2V_OK;
    int dev = 0, sms = 0;
    C2V_CUDA_OK(cudaGetDevice(&dev));
    C2V_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    C2V_CUDA_OK(cudaFuncSetAttribute(backward_dc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dct::SMEM_BYTES));
    int grid = sms / n_db;                                 // one CTA per SM over the d windows
    if (grid > a.n_tiles) grid = a.n_tiles;
    if (grid < 1) grid = 1;
    backward_dc_tc_kernel<<<dim3(grid, n_db), dct::THREADS, dct::SMEM_BYTES, st>>>(a, dx, dx_absmax, img, hdr, g_emb_t, g_emb_p,
                                                                                sv_mask & 7, n_hb);
    C2V_LAUNCH_OK("backward_dc_tc_kernel");
    return C2V_OK;
}

}  // namespace c2v
