// c2v_encode_tma.cu -- K1c: fused gather + encode + attention with the gathers done by the TMA
// (`cp.async.bulk.tensor.2d ... tile::gather4`), for terminal_embed = path_embed = encode = 128.
//
// Why: in K1b (c2v_encode_tcgen05.cu) the SM issue slots are the scarcest resource -- two thirds of
// all warp instructions are the producers' address arithmetic, LDG, index shuffles and register
// double-buffering around ~1.5 us of HBM latency.  Here one warp issues 32 gather4 operations per
// 64-wide k-block (4 embedding half-rows each, indices straight from the int64 arrays) into a raw
// fp32 smem stage; the latency lives in the TMA queue, and 8 converter warps only do
// LDS.128 -> hi/lo fp16 split -> STS.64 into the UMMA K-major SWIZZLE_128B tiles.
// Numerics, MMA schedule and epilogue are those of K1b (model.py:48-69 + 90-96; 3-pass fp16 split).
//
// Warps (20):  0-7 epilogue | 8-15 converters | 16 MMA issuer | 17 W producer | 18 gather | 19 TMEM alloc
// smem (197 KB): 2 raw stages [128 rows x 64 fp32] (64 KB) | 2 operand stages {A_hi,A_lo,W_hi,W_lo}
//                (128 KB) | gamma'/beta'/attn | LN exchange | mbarriers
#include <cuda.h>

#include <cstdlib>

#include "c2v_tc_epilogue.cuh"

namespace c2v {

namespace tm {
constexpr int ROWS = tce::ROWS, H = tce::H, E = 128, D = 3 * E;
constexpr int KB = 64, NKB = D / KB;                  // 6 k-blocks per tile
constexpr int RAW_STAGES = 2, OP_STAGES = 2;
constexpr int RAW_BYTES = ROWS * KB * 4;              // 32 KB
constexpr int TILE_BYTES = ROWS * KB * 2;             // 16 KB fp16 tile
constexpr int OP_BYTES = 4 * TILE_BYTES;              // 64 KB
constexpr int W_KB_BYTES = 2 * TILE_BYTES;
constexpr int N_CONV_WARPS = 8;
constexpr int CONV_WARP0 = tce::N_EPI_WARPS;          // 8
constexpr int MISC_WARP0 = CONV_WARP0 + N_CONV_WARPS; // 16
constexpr int THREADS = (MISC_WARP0 + 4) * 32;        // 640
constexpr int ROWS_PER_CW = ROWS / N_CONV_WARPS;      // 16 rows per converter warp
constexpr int LDS_PER_ITEM = ROWS_PER_CW / 2;         // 8 x LDS.128 (two 256-B half rows per instruction)
constexpr int TMEM_COLS = 256;
constexpr int SMEM_RAW_OFF = 0;
constexpr int SMEM_OP_OFF = RAW_STAGES * RAW_BYTES;
constexpr int SMEM_VEC_OFF = SMEM_OP_OFF + OP_STAGES * OP_BYTES;
constexpr int SMEM_XCH_OFF = SMEM_VEC_OFF + tce::VEC_BYTES;
constexpr int SMEM_BAR_OFF = SMEM_XCH_OFF + tce::XCH_BYTES;
constexpr int SMEM_BYTES = SMEM_BAR_OFF + 128 + 1024;
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(ROWS >> 4) << 24);
static_assert(NKB % RAW_STAGES == 0 && NKB % OP_STAGES == 0, "stage index is a function of kb only");
}  // namespace tm

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2,
                                            int r3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
                 "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
// L2 prefetch of 4 half rows: the HBM latency of the NEXT tile's gathers is paid here, without
// holding shared memory; the real gather4 then hits L2.
__device__ __forceinline__ void tma_prefetch_gather4(const CUtensorMap *map, int col, int r0, int r1, int r2, int r3) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile::gather4 [%0, {%1, %2, %3, %4, %5}];"
                 ::"l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <bool DROPOUT>
__global__ void __launch_bounds__(tm::THREADS, 1)
encode_tma_kernel(const EncodeArgs a, const __grid_constant__ CUtensorMap map_t, const __grid_constant__ CUtensorMap map_p)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char *smem = smem_raw + (base - raw);
    float *s_vec = reinterpret_cast<float *>(smem + tm::SMEM_VEC_OFF);
    float *s_xch = reinterpret_cast<float *>(smem + tm::SMEM_XCH_OFF);
    const uint32_t bar_base = base + tm::SMEM_BAR_OFF;
    // 8-byte barriers: raw_full[2] @0, raw_empty[2] @16, op_full[2] @32, op_empty[2] @48,
    //                  tmem_full[2] @64, tmem_empty[2] @80, tmem ptr @96
    const uint32_t bar_rfull = bar_base, bar_rempty = bar_base + 16, bar_ofull = bar_base + 32,
                   bar_oempty = bar_base + 48, bar_tfull = bar_base + 64, bar_tempty = bar_base + 80;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + tm::SMEM_BAR_OFF + 96);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    long long *status = a.ws.status;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_rfull + 8 * s, 1);
            mbar_init(bar_rempty + 8 * s, tm::N_CONV_WARPS);
            mbar_init(bar_ofull + 8 * s, tm::N_CONV_WARPS + 1);
            mbar_init(bar_oempty + 8 * s, 1);
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, tce::N_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == tm::MISC_WARP0 + 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"((uint32_t)tm::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tce_fill_vectors(a, s_vec, tid);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // item = (tile tl, k-block kb); both rings have 2 stages and NKB = 6 is even, so the stage of an
    // item is kb & 1 and its use count is tl*3 + kb/2  ->  phase parity (tl*3 + kb/2) & 1.
    if (warp < tce::N_EPI_WARPS) {
        // =============================== EPILOGUE ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
        tce_epilogue_loop<DROPOUT>(a, s_vec, s_xch, tmem_base, bar_tfull, bar_tempty, warp, lane, my_tiles, status);
    } else if (warp < tm::MISC_WARP0) {
        // =============================== CONVERTERS ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        const int cw = warp - tm::CONV_WARP0;         // rows 16*cw .. 16*cw+15 of every tile
        const int sub_row = lane >> 4, q = lane & 15;
        uint32_t ld_off[tm::LDS_PER_ITEM], st_off[tm::LDS_PER_ITEM];
#pragma unroll
        for (int j = 0; j < tm::LDS_PER_ITEM; ++j) {
            const int r = cw * tm::ROWS_PER_CW + 2 * j + sub_row;
            ld_off[j] = (uint32_t)(r * (tm::KB * 4) + q * 16);
            st_off[j] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((q >> 1) ^ (r & 7)) & 7) << 4) + (q & 1) * 8);
        }
        for (int tl = 0; tl < my_tiles; ++tl) {
#pragma unroll 2
            for (int kb = 0; kb < tm::NKB; ++kb) {
                const int st = kb & 1;
                const uint32_t phase = (uint32_t)(tl * 3 + (kb >> 1)) & 1u;
                const uint32_t rawb = base + tm::SMEM_RAW_OFF + st * tm::RAW_BYTES;
                const uint32_t a_hi = base + tm::SMEM_OP_OFF + st * tm::OP_BYTES, a_lo = a_hi + tm::TILE_BYTES;
                mbar_wait(bar_rfull + 8 * st, phase, status);             // gathered fp32 rows have landed
                float4 v[tm::LDS_PER_ITEM];
#pragma unroll
                for (int j = 0; j < tm::LDS_PER_ITEM; ++j) v[j] = lds_v4(rawb + ld_off[j]);
                // the asm volatile loads above complete in order before this arrive: the raw stage can be
                // refilled by the next gather while this warp converts out of registers
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_rempty + 8 * st);
                mbar_wait(bar_oempty + 8 * st, phase ^ 1u, status);       // MMAs of the previous use retired
#pragma unroll
                for (int j = 0; j < tm::LDS_PER_ITEM; ++j) {
                    const __half2 h01 = __floats2half2_rn(v[j].x, v[j].y), h23 = __floats2half2_rn(v[j].z, v[j].w);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    const __half2 l01 = __floats2half2_rn(v[j].x - f01.x, v[j].y - f01.y);
                    const __half2 l23 = __floats2half2_rn(v[j].z - f23.x, v[j].w - f23.y);
                    sts_v2(a_hi + st_off[j], pack_h2(h01), pack_h2(h23));
                    sts_v2(a_lo + st_off[j], pack_h2(l01), pack_h2(l23));
                }
                fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_ofull + 8 * st);
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == tm::MISC_WARP0) {
            // =============================== MMA ISSUER ===============================
            if (lane == 0) {
                for (int tl = 0; tl < my_tiles; ++tl) {
                    const int acc = tl & 1;
                    const uint32_t acc_phase = (uint32_t)(tl >> 1) & 1u;
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u, status);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * tm::H);
#pragma unroll 1
                    for (int kb = 0; kb < tm::NKB; ++kb) {
                        const int st = kb & 1;
                        const uint32_t phase = (uint32_t)(tl * 3 + (kb >> 1)) & 1u;
                        mbar_wait(bar_ofull + 8 * st, phase, status);
                        tc_fence_after();
                        const uint32_t sa = base + tm::SMEM_OP_OFF + st * tm::OP_BYTES;
#pragma unroll
                        for (int k = 0; k < tm::KB / 16; ++k) {
                            const uint64_t a_hi = umma_desc(sa + k * 32);
                            const uint64_t a_lo = umma_desc(sa + tm::TILE_BYTES + k * 32);
                            const uint64_t w_hi = umma_desc(sa + 2 * tm::TILE_BYTES + k * 32);
                            const uint64_t w_lo = umma_desc(sa + 3 * tm::TILE_BYTES + k * 32);
                            umma_f16(d_tmem, a_hi, w_hi, tm::IDESC, (kb | k) != 0 ? 1u : 0u);
                            umma_f16(d_tmem, a_lo, w_hi, tm::IDESC, 1u);
                            umma_f16(d_tmem, a_hi, w_lo, tm::IDESC, 1u);
                        }
                        umma_commit(bar_oempty + 8 * st);
                    }
                    umma_commit(bar_tfull + 8 * acc);
                }
            }
            __syncwarp();
        } else if (warp == tm::MISC_WARP0 + 1) {
            // =============================== W PRODUCER ===============================
            if (lane == 0) {
                const uint8_t *img = reinterpret_cast<const uint8_t *>(a.ws.w_hi);
                for (int tl = 0; tl < my_tiles; ++tl)
#pragma unroll 1
                    for (int kb = 0; kb < tm::NKB; ++kb) {
                        const int st = kb & 1;
                        const uint32_t phase = (uint32_t)(tl * 3 + (kb >> 1)) & 1u;
                        mbar_wait(bar_oempty + 8 * st, phase ^ 1u, status);
                        mbar_arrive_expect_tx(bar_ofull + 8 * st, tm::W_KB_BYTES);
                        bulk_copy_g2s(base + tm::SMEM_OP_OFF + st * tm::OP_BYTES + 2 * tm::TILE_BYTES,
                                      img + (size_t)kb * tm::W_KB_BYTES, tm::W_KB_BYTES, bar_ofull + 8 * st);
                    }
            }
            __syncwarp();
        } else if (warp == tm::MISC_WARP0 + 2) {
            // =============================== GATHER (TMA) ===============================
            // lane l owns rows 4l .. 4l+3 of every tile: one gather4 per k-block per lane.
            int idx[3][4];                         // [start|path|end][4 rows], int32 row coordinates
            auto load_idx = [&](int tl, int (&out)[3][4]) {
                const long long row0 = ((long long)blockIdx.x + (long long)tl * gridDim.x) * tm::ROWS + 4 * lane;
                int bad = 0;
#pragma unroll
                for (int w = 0; w < 3; ++w) {
                    const long long *src = w == 0 ? a.starts : (w == 1 ? a.paths : a.ends);
                    const long long lim = w == 1 ? a.P : a.T;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        long long v = 0;
                        if (tl < my_tiles && row0 + j < a.N) v = src[row0 + j];
                        if (v < 0 || v >= lim) { v = 0; ++bad; }
                        out[w][j] = (int)v;
                    }
                }
                if (bad) atomicAdd((unsigned long long *)status, (unsigned long long)bad);
            };
            int nxt[3][4];
            load_idx(0, idx);
            for (int tl = 0; tl < my_tiles; ++tl) {
                load_idx(tl + 1, nxt);             // the next tile's indices ...
                if (tl + 1 < my_tiles) {           // ... and its rows into L2 (6 x 4 half rows per lane)
#pragma unroll
                    for (int kb = 0; kb < tm::NKB; ++kb)
                        tma_prefetch_gather4((kb >> 1) == 1 ? &map_p : &map_t, (kb & 1) * tm::KB,
                                             nxt[kb >> 1][0], nxt[kb >> 1][1], nxt[kb >> 1][2], nxt[kb >> 1][3]);
                }
#pragma unroll
                for (int kb = 0; kb < tm::NKB; ++kb) {
                    const int st = kb & 1, sub = kb >> 1;
                    const uint32_t phase = (uint32_t)(tl * 3 + (kb >> 1)) & 1u;
                    mbar_wait(bar_rempty + 8 * st, phase ^ 1u, status);
                    if (lane == 0) mbar_arrive_expect_tx(bar_rfull + 8 * st, tm::RAW_BYTES);
                    __syncwarp();
                    tma_gather4(base + tm::SMEM_RAW_OFF + st * tm::RAW_BYTES + lane * 4 * (tm::KB * 4),
                                sub == 1 ? &map_p : &map_t, (kb & 1) * tm::KB,
                                idx[sub][0], idx[sub][1], idx[sub][2], idx[sub][3], bar_rfull + 8 * st);
                }
#pragma unroll
                for (int w = 0; w < 3; ++w)
#pragma unroll
                    for (int j = 0; j < 4; ++j) idx[w][j] = nxt[w][j];
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == tm::MISC_WARP0 + 3) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)tm::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------
// host: tensor maps (driver entry point fetched through the runtime: no libcuda link dependency)
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, 128] fp32 table, box = {64 columns, 1 row}: tile::gather4 fetches 4 such rows per instruction
static int make_table_map(CUtensorMap *m, const float *table, long long rows)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return C2V_EUNSUPPORTED; }
    const cuuint64_t gdim[2] = {(cuuint64_t)tm::E, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)tm::E * 4};
    const cuuint32_t box[2] = {(cuuint32_t)tm::KB, 1};
    const cuuint32_t estride[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(table), gdim, gstride, box, estride,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return C2V_ECUDA; }
    return C2V_OK;
}

bool encode_tma_available() { return get_encode_fn() != nullptr; }

int launch_encode_tma(const EncodeArgs &a, cudaStream_t st)
{
    CUtensorMap map_t, map_p;
    int rc = make_table_map(&map_t, a.emb_t, a.T);
    if (rc != C2V_OK) return rc;
    rc = make_table_map(&map_p, a.emb_p, a.P);
    if (rc != C2V_OK) return rc;
    int dev = 0, sms = 0;
    C2V_CUDA_OK(cudaGetDevice(&dev));
    C2V_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    auto kern = a.drop_p > 0.0f ? encode_tma_kernel<true> : encode_tma_kernel<false>;
    C2V_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tm::SMEM_BYTES));
    int grid = a.n_tiles < sms ? a.n_tiles : sms;
    if (grid < 1) grid = 1;
    kern<<<grid, tm::THREADS, tm::SMEM_BYTES, st>>>(a, map_t, map_p);
    C2V_LAUNCH_OK("encode_tma_kernel");
    return C2V_OK;
}

}  // namespace c2v
