// c2v_encode_cpa.cu -- K1d: fused gather + encode + attention with asynchronous-copy loaders.
//
// Same numerics / MMA schedule / epilogue as K1b (c2v_encode_tcgen05.cu; model.py:48-69 + 90-96, 3-pass fp16
// split), different producer side.  In K1b the issue slots are the scarcest resource: two thirds of all warp
// instructions are the producers' LDG + address arithmetic + register double-buffering + conversions.  Here
//   * 4 loader warps copy the fp32 half rows global -> shared with cp.async (LDGSTS.128, no registers, no
//     scoreboard waits; completion reported to an mbarrier by cp.async.mbarrier.arrive.noinc), and
//   * 8 converter warps do LDS.128 -> hi/lo fp16 split -> STS.64 into the UMMA K-major SWIZZLE_128B tiles,
// so the HBM latency lives in the LSU's async-copy queue instead of in registers.
//
// Warps (24):  0-7 epilogue | 8-15 converters | 16-19 loaders | 20 MMA issuer | 21 W producer | 22 TMEM alloc
// smem (197 KB): 2 raw stages [128 rows x 64 fp32] (64 KB) | 2 operand stages {A_hi,A_lo,W_hi,W_lo}
//                (128 KB) | gamma'/beta'/attn | LN exchange | mbarriers

#include <cstdlib>

#include "c2v_tc_epilogue.cuh"

namespace c2v {

namespace ca {
constexpr int ROWS = tce::ROWS, H = tce::H, E = 128, D = 3 * E;
constexpr int KB = 64, NKB = D / KB;                  // 6 k-blocks per tile
constexpr int RAW_STAGES = 2, OP_STAGES = 2;
constexpr int RAW_BYTES = ROWS * KB * 4;              // 32 KB
constexpr int TILE_BYTES = ROWS * KB * 2;             // 16 KB fp16 tile
constexpr int OP_BYTES = 4 * TILE_BYTES;              // 64 KB
constexpr int W_KB_BYTES = 2 * TILE_BYTES;
constexpr int N_CONV_WARPS = 8;
constexpr int CONV_WARP0 = tce::N_EPI_WARPS;          // 8
constexpr int N_LOAD_WARPS = 4;
constexpr int LOAD_WARP0 = CONV_WARP0 + N_CONV_WARPS; // 16
constexpr int MISC_WARP0 = LOAD_WARP0 + N_LOAD_WARPS; // 20
constexpr int THREADS = (MISC_WARP0 + 4) * 32;        // 768
constexpr int ROWS_PER_LW = ROWS / N_LOAD_WARPS;      // 32 rows of every tile per loader warp
constexpr int CPA_PER_ITEM = ROWS_PER_LW / 2;         // 16 x LDGSTS.128 (two 256-B half rows per instruction)
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
}  // namespace ca

__device__ __forceinline__ void cp_async_cg16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the mbarrier receives one arrival from this thread once all of its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <bool DROPOUT>
__global__ void __launch_bounds__(ca::THREADS, 1)
encode_cpa_kernel(const EncodeArgs a)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char *smem = smem_raw + (base - raw);
    float *s_vec = reinterpret_cast<float *>(smem + ca::SMEM_VEC_OFF);
    float *s_xch = reinterpret_cast<float *>(smem + ca::SMEM_XCH_OFF);
    const uint32_t bar_base = base + ca::SMEM_BAR_OFF;
    // 8-byte barriers: raw_full[2] @0, raw_empty[2] @16, op_full[2] @32, op_empty[2] @48,
    //                  tmem_full[2] @64, tmem_empty[2] @80, tmem ptr @96
    const uint32_t bar_rfull = bar_base, bar_rempty = bar_base + 16, bar_ofull = bar_base + 32,
                   bar_oempty = bar_base + 48, bar_tfull = bar_base + 64, bar_tempty = bar_base + 80;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + ca::SMEM_BAR_OFF + 96);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int my_tiles = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    long long *status = a.ws.status;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_rfull + 8 * s, ca::N_LOAD_WARPS * 32);
            mbar_init(bar_rempty + 8 * s, ca::N_CONV_WARPS);
            mbar_init(bar_ofull + 8 * s, ca::N_CONV_WARPS + 1);
            mbar_init(bar_oempty + 8 * s, 1);
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, tce::N_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == ca::MISC_WARP0 + 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"((uint32_t)ca::TMEM_COLS) : "memory");
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
    } else if (warp < ca::LOAD_WARP0) {
        // =============================== CONVERTERS ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        const int cw = warp - ca::CONV_WARP0;         // rows 16*cw .. 16*cw+15 of every tile
        const int sub_row = lane >> 4, q = lane & 15;
        uint32_t ld_off[ca::LDS_PER_ITEM], st_off[ca::LDS_PER_ITEM];
#pragma unroll
        for (int j = 0; j < ca::LDS_PER_ITEM; ++j) {
            const int r = cw * ca::ROWS_PER_CW + 2 * j + sub_row;
            ld_off[j] = (uint32_t)(r * (ca::KB * 4) + q * 16);
            st_off[j] = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((q >> 1) ^ (r & 7)) & 7) << 4) + (q & 1) * 8);
        }
        for (int tl = 0; tl < my_tiles; ++tl) {
#pragma unroll 2
            for (int kb = 0; kb < ca::NKB; ++kb) {
                const int st = kb & 1;
                const uint32_t phase = (uint32_t)(tl * 3 + (kb >> 1)) & 1u;
                const uint32_t rawb = base + ca::SMEM_RAW_OFF + st * ca::RAW_BYTES;
                const uint32_t a_hi = base + ca::SMEM_OP_OFF + st * ca::OP_BYTES, a_lo = a_hi + ca::TILE_BYTES;
                mbar_wait(bar_rfull + 8 * st, phase, status);             // gathered fp32 rows have landed
                float4 v[ca::LDS_PER_ITEM];
#pragma unroll
                for (int j = 0; j < ca::LDS_PER_ITEM; ++j) v[j] = lds_v4(rawb + ld_off[j]);
                // the asm volatile loads above complete in order before this arrive: the raw stage can be
                // refilled by the next gather while this warp converts out of registers
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_rempty + 8 * st);
                mbar_wait(bar_oempty + 8 * st, phase ^ 1u, status);       // MMAs of the previous use retired
                if (!C2V_EXPT(a.flags, 32))         // (timing experiment: skip the conversion + stores)
#pragma unroll
                for (int j = 0; j < ca::LDS_PER_ITEM; ++j) {
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
    } else if (warp < ca::MISC_WARP0) {
        // =============================== LOADERS (cp.async) ===============================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const int lw = warp - ca::LOAD_WARP0;        // rows 32*lw .. 32*lw+31 of every tile; lane l owns row 32*lw+l's indices
        const int sub_row = lane >> 4, q = lane & 15;
        const char *tab_t = reinterpret_cast<const char *>(a.emb_t) + q * 16;
        const char *tab_p = reinterpret_cast<const char *>(a.emb_p) + q * 16;
        const uint32_t dst_lane = (uint32_t)((lw * ca::ROWS_PER_LW + sub_row) * (ca::KB * 4) + q * 16);
        long long rs = 0, rp = 0, re = 0;            // raw indices of the NEXT tile (prefetched)
        uint32_t off_s = 0, off_p = 0, off_e = 0;    // byte offsets of this lane's row in the tables
        auto fetch_idx = [&](int tl) {
            rs = rp = re = 0;
            if (tl < my_tiles) {
                const long long row = ((long long)blockIdx.x + (long long)tl * gridDim.x) * ca::ROWS + lw * ca::ROWS_PER_LW + lane;
                if (row < a.N) { rs = a.starts[row]; rp = a.paths[row]; re = a.ends[row]; }
            }
        };
        auto adopt_idx = [&]() {
            int bad = 0;
            if (rs < 0 || rs >= a.T) { rs = 0; ++bad; }
            if (rp < 0 || rp >= a.P) { rp = 0; ++bad; }
            if (re < 0 || re >= a.T) { re = 0; ++bad; }
            if (bad) atomicAdd((unsigned long long *)status, (unsigned long long)bad);
            off_s = (uint32_t)rs * (ca::E * 4); off_p = (uint32_t)rp * (ca::E * 4); off_e = (uint32_t)re * (ca::E * 4);
        };
        fetch_idx(0);
        for (int tl = 0; tl < my_tiles; ++tl) {
            adopt_idx();
            fetch_idx(tl + 1);
#pragma unroll
            for (int kb = 0; kb < ca::NKB; ++kb) {
                const int st = kb & 1, sub = kb >> 1;
                const uint32_t phase = (uint32_t)(tl * 3 + (kb >> 1)) & 1u;
                const char *tab = (sub == 1 ? tab_p : tab_t) + (kb & 1) * (ca::KB * 4);
                const uint32_t off = sub == 0 ? off_s : (sub == 1 ? off_p : off_e);
                const uint32_t dst = base + ca::SMEM_RAW_OFF + st * ca::RAW_BYTES + dst_lane;
                mbar_wait(bar_rempty + 8 * st, phase ^ 1u, status);
#pragma unroll
                for (int j = 0; j < ca::CPA_PER_ITEM; ++j) {
                    const uint32_t o = __shfl_sync(0xffffffffu, off, 2 * j + sub_row);
                    cp_async_cg16(dst + j * 2 * (ca::KB * 4), tab + o);
                }
                cp_async_mbar_arrive_noinc(bar_rfull + 8 * st);
                if (C2V_EXPT(a.flags, 2)) {      // (experiment, off by default: measured slower)
                    // Only two raw stages (64 KB) can be in flight per SM, which at HBM latency caps the kernel
                    // (profiles/README.md).  Pull the half rows of item +2 -- the one that cannot be issued yet --
                    // towards L2 now: this lane's own row, two 128-B lines.
                    const char *pf;
                    if (kb + 2 < ca::NKB) {
                        const int sub2 = (kb + 2) >> 1;
                        pf = (sub2 == 1 ? reinterpret_cast<const char *>(a.emb_p) : reinterpret_cast<const char *>(a.emb_t)) +
                             (sub2 == 0 ? off_s : (sub2 == 1 ? off_p : off_e)) + (kb & 1) * (ca::KB * 4);
                    } else {                                   // first sub-vector (starts) of the next tile
                        long long r2 = rs;
                        if (r2 < 0 || r2 >= a.T) r2 = 0;
                        pf = reinterpret_cast<const char *>(a.emb_t) + r2 * (ca::E * 4) + (kb & 1) * (ca::KB * 4);
                    }
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 128));
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == ca::MISC_WARP0) {
            // =============================== MMA ISSUER ===============================
            if (lane == 0) {
                for (int tl = 0; tl < my_tiles; ++tl) {
                    const int acc = tl & 1;
                    const uint32_t acc_phase = (uint32_t)(tl >> 1) & 1u;
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1u, status);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ca::H);
#pragma unroll 1
                    for (int kb = 0; kb < ca::NKB; ++kb) {
                        const int st = kb & 1;
                        const uint32_t phase = (uint32_t)(tl * 3 + (kb >> 1)) & 1u;
                        mbar_wait(bar_ofull + 8 * st, phase, status);
                        tc_fence_after();
                        const uint32_t sa = base + ca::SMEM_OP_OFF + st * ca::OP_BYTES;
#pragma unroll
                        for (int k = 0; k < ca::KB / 16; ++k) {
                            const uint64_t a_hi = umma_desc(sa + k * 32);
                            const uint64_t a_lo = umma_desc(sa + ca::TILE_BYTES + k * 32);
                            const uint64_t w_hi = umma_desc(sa + 2 * ca::TILE_BYTES + k * 32);
                            const uint64_t w_lo = umma_desc(sa + 3 * ca::TILE_BYTES + k * 32);
                            umma_f16(d_tmem, a_hi, w_hi, ca::IDESC, (kb | k) != 0 ? 1u : 0u);
                            if (!C2V_EXPT(a.flags, 4)) umma_f16(d_tmem, a_lo, w_hi, ca::IDESC, 1u);
                            if (!C2V_EXPT(a.flags, 8)) umma_f16(d_tmem, a_hi, w_lo, ca::IDESC, 1u);
                        }
                        umma_commit(bar_oempty + 8 * st);
                    }
                    umma_commit(bar_tfull + 8 * acc);
                }
            }
            __syncwarp();
        } else if (warp == ca::MISC_WARP0 + 1) {
            // =============================== W PRODUCER ===============================
            if (lane == 0) {
                const uint8_t *img = reinterpret_cast<const uint8_t *>(a.ws.w_hi);
                for (int tl = 0; tl < my_tiles; ++tl)
#pragma unroll 1
                    for (int kb = 0; kb < ca::NKB; ++kb) {
                        const int st = kb & 1;
                        const uint32_t phase = (uint32_t)(tl * 3 + (kb >> 1)) & 1u;
                        mbar_wait(bar_oempty + 8 * st, phase ^ 1u, status);
                        mbar_arrive_expect_tx(bar_ofull + 8 * st, ca::W_KB_BYTES);
                        bulk_copy_g2s(base + ca::SMEM_OP_OFF + st * ca::OP_BYTES + 2 * ca::TILE_BYTES,
                                      img + (size_t)kb * ca::W_KB_BYTES, ca::W_KB_BYTES, bar_ofull + 8 * st);
                    }
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == ca::MISC_WARP0 + 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)ca::TMEM_COLS) : "memory");
    }
}

int launch_encode_cpa(const EncodeArgs &a, cudaStream_t st)
{
    int dev = 0, sms = 0;
    C2V_CUDA_OK(cudaGetDevice(&dev));
    C2V_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    auto kern = a.drop_p > 0.0f ? encode_cpa_kernel<true> : encode_cpa_kernel<false>;
    EncodeArgs b = a;
    const char *dbg = getenv("C2V_DEBUG_FLAGS");      // timing experiments only (results become wrong)
    if (dbg) b.flags |= atoi(dbg);
    C2V_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ca::SMEM_BYTES));
    int grid = a.n_tiles < sms ? a.n_tiles : sms;
    if (grid < 1) grid = 1;
    kern<<<grid, ca::THREADS, ca::SMEM_BYTES, st>>>(b);
    C2V_LAUNCH_OK("encode_cpa_kernel");
    return C2V_OK;
}

}  // namespace c2v
