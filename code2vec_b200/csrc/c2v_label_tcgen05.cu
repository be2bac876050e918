// c2v_label_tcgen05.cu -- K2: label logits  outputs = cv . W_out^T + b  (model.py:83) on tcgen05.
//
// Same fp32-accurate scheme as the encode kernel: both operands are split into fp16 hi + lo and
// three kind::f16 MMAs (hi.hi + lo.hi + hi.lo) accumulate in fp32 in TMEM.  W_out is scaled by a
// power of two (from its absmax) before the split; the epilogue multiplies by the exact inverse
// and adds the bias.  Plain TF32 fails the 1e-4 bar on trained weights.
//
// Three launches: absmax(W_out) -> split cv and W_out into UMMA K-major SWIZZLE_128B tile images
// ([128 rows x 64 k] fp16, hi then lo, per k-block) -> one CTA per 128x128 output tile: two 64 KB
// cp.async.bulk copies, 24 tcgen05.mma, TMEM -> registers -> padded smem -> coalesced 512-B row
// stores with the bias added.  Output-write bound (B*C*4 bytes).
#include <cuda_fp16.h>

#include "c2v_common.cuh"

namespace c2v {

namespace lt {
constexpr int TM = 128, TN = 128, KB = 64;
constexpr int TILE_BYTES = 128 * KB * 2;          // 16 KB
constexpr int MAX_NKB = 2;                        // encode_size 64 or 128
constexpr int STAGE_LD = 132;                     // padded fp32 row of the output staging tile
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}  // namespace lt

__host__ __device__ __forceinline__ uint32_t lt_sw128(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}
__device__ __forceinline__ uint32_t lt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void absmax_kernel(const float *__restrict__ x, long long n, unsigned *__restrict__ out_bits)
{
    float m = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order as uints
}

// X [R, K] fp32 row-major -> image: for each 128-row tile, for each k-block: {hi tile, lo tile}.
// scale_bits == nullptr: no scaling.  Rows >= R are zero-filled.  hdr[0] = 1/scale, hdr[1] = scale.
__global__ void split_rows_kernel(const float *__restrict__ X, long long R, int K, int nkb,
                                  const unsigned *__restrict__ scale_bits, uint8_t *__restrict__ img,
                                  float *__restrict__ hdr)
{
    float scale = 1.0f;
    if (scale_bits) {
        const float mx = __uint_as_float(*scale_bits);
        if (mx > 0.0f && mx < 3.0e38f) {
            int e;
            frexpf(mx, &e);
            int k = 14 - e;
            k = k > 60 ? 60 : (k < -60 ? -60 : k);
            scale = ldexpf(1.0f, k);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) { hdr[0] = 1.0f / scale; hdr[1] = scale; }
    }
    const long long tiles = (R + 127) / 128;
    const long long total = tiles * 128 * (long long)(nkb * 16);          // float4 groups
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const int k4 = (int)(g % (nkb * 16));
        const long long row = g / (nkb * 16);
        const int k = k4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < R && k < K) v = *reinterpret_cast<const float4 *>(X + row * K + k);
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
        const long long tile = row >> 7;
        const int r = (int)(row & 127), kb = k / lt::KB, kk = k % lt::KB;
        uint8_t *base = img + (tile * nkb + kb) * (size_t)(2 * lt::TILE_BYTES);
        const uint32_t off = lt_sw128(r, kk);
        *reinterpret_cast<uint2 *>(base + off) = make_uint2(*reinterpret_cast<const unsigned *>(&h01), *reinterpret_cast<const unsigned *>(&h23));
        *reinterpret_cast<uint2 *>(base + lt::TILE_BYTES + off) = make_uint2(*reinterpret_cast<const unsigned *>(&l01), *reinterpret_cast<const unsigned *>(&l23));
    }
}

__device__ __forceinline__ void lt_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (unsigned spins = 0; !ok; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && spins > (1u << 26)) __trap();
    }
}

// Persistent label GEMM.  Grid = (G, n_mt): CTA (g, mt) keeps the cv image of m-tile mt resident and walks
// the n-tiles g, g+G, ...; W_out tiles are double buffered in smem, accumulators in TMEM, so the bulk copy of
// tile i+1, the MMAs of tile i and the epilogue (TMEM -> registers -> +bias -> row stores, running argmax) of
// tile i-1 overlap.  Warps 0-3: epilogue (one output row per thread); warp 4: loads + TMEM alloc; warp 5: MMA.
__global__ void __launch_bounds__(192, 1)
label_gemm_tcgen05_kernel(const uint8_t *__restrict__ imgA, const uint8_t *__restrict__ imgB,
                          const float *__restrict__ bias, const float *__restrict__ hdr,
                          float *__restrict__ out, int M, long long N, int nkb, int n_nt)
{
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = lt_smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char *smem = smem_raw + (base - raw);
    const int op_bytes = nkb * 2 * lt::TILE_BYTES;                    // one operand tile image: 32 KB per k-block
    const uint32_t sA = base, sB0 = base + op_bytes;
    constexpr int STAGE_BYTES = 4 * 32 * lt::STAGE_LD * 4;           // per-warp [32 rows][132] fp32 output staging
    float *stage_all = reinterpret_cast<float *>(smem + 2 * op_bytes);
    const int bar_off = 2 * op_bytes + STAGE_BYTES;
    const uint32_t bars = base + bar_off;                             // a_full @0, b_full @8, b_empty @24,
    const uint32_t bar_afull = bars, bar_bfull = bars + 8, bar_bempty = bars + 24,   // t_full[2] @40, t_empty[2] @56
                   bar_tfull = bars + 40, bar_tempty = bars + 56;
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(smem + bar_off + 72);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mt = blockIdx.y, g = blockIdx.x, G = gridDim.x;
    const int my_tiles = g < n_nt ? (n_nt - g + G - 1) / G : 0;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_afull));
        for (int s = 0; s < 2; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_bfull + 8 * s));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_bempty + 8 * s));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_tfull + 8 * s));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(bar_tempty + 8 * s));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(lt_smem_u32(tmem_ptr_smem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_ptr_smem;

    if (warp == 4) {
        if (lane == 0 && my_tiles > 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_afull), "r"((uint32_t)op_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sA), "l"(imgA + (size_t)mt * op_bytes), "r"((uint32_t)op_bytes), "r"(bar_afull) : "memory");
            for (int i = 0; i < my_tiles; ++i) {        // single W_out buffer: refilled as soon as tile i-1's MMAs retire,
                const long long nt = g + (long long)i * G;   // i.e. while the epilogue of tile i-1 is still storing
                lt_mbar_wait(bar_bempty, ((uint32_t)i & 1u) ^ 1u);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_bfull), "r"((uint32_t)op_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sB0), "l"(imgB + (size_t)nt * op_bytes), "r"((uint32_t)op_bytes), "r"(bar_bfull) : "memory");
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        if (lane == 0 && my_tiles > 0) {
            lt_mbar_wait(bar_afull, 0);
            for (int i = 0; i < my_tiles; ++i) {
                const int st = i & 1;
                const uint32_t ph = (uint32_t)(i >> 1) & 1u;
                lt_mbar_wait(bar_tempty + 8 * st, ph ^ 1u);
                lt_mbar_wait(bar_bfull, (uint32_t)i & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem + (uint32_t)(st * lt::TN);
                for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                    for (int k = 0; k < lt::KB / 16; ++k) {
                        auto desc = [](uint32_t a) {
                            return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
                        };
                        const uint32_t a0 = sA + kb * 2 * lt::TILE_BYTES + k * 32;
                        const uint32_t b0 = sB0 + kb * 2 * lt::TILE_BYTES + k * 32;
                        const uint64_t a_hi = desc(a0), a_lo = desc(a0 + lt::TILE_BYTES), b_hi = desc(b0), b_lo = desc(b0 + lt::TILE_BYTES);
                        auto mma = [&](uint64_t ad, uint64_t bd, uint32_t acc) {
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(d_tmem), "l"(ad), "l"(bd), "r"(lt::IDESC), "r"(acc) : "memory");
                        };
                        mma(a_hi, b_hi, (kb | k) != 0 ? 1u : 0u);
                        mma(a_lo, b_hi, 1u);
                        mma(a_hi, b_lo, 1u);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_bempty) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_tfull + 8 * st) : "memory");
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue: thread = output row of the m-tile; rows leave through a per-warp padded smem tile
        //      so that every global store instruction writes one full 512-B row segment
        const float inv_scale = hdr[0];
        const bool vec_ok = (N % 4 == 0);
        float *stg = stage_all + warp * 32 * lt::STAGE_LD;
        for (int i = 0; i < my_tiles; ++i) {
            const int st = i & 1;
            const long long nt = g + (long long)i * G;
            lt_mbar_wait(bar_tfull + 8 * st, (uint32_t)(i >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(st * lt::TN);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                               "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(taddr + c * 32));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == 3) {   // all 128 columns have left TMEM: the accumulator can be reused
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tempty + 8 * st) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(stg + lane * lt::STAGE_LD + c * 32 + j) =
                        make_float4(__uint_as_float(r[j]) * inv_scale, __uint_as_float(r[j + 1]) * inv_scale,
                                    __uint_as_float(r[j + 2]) * inv_scale, __uint_as_float(r[j + 3]) * inv_scale);
            }
            __syncwarp();
            // coalesced copy-out (+bias, model.py:83): each store instruction writes one full 512-B row segment
            const long long col = nt * lt::TN + lane * 4;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) {
                if (col + 3 < N) bv = __ldg(reinterpret_cast<const float4 *>(bias + col));
                else { if (col < N) bv.x = bias[col]; if (col + 1 < N) bv.y = bias[col + 1]; if (col + 2 < N) bv.z = bias[col + 2]; }
            }
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
                const int orow_i = mt * lt::TM + warp * 32 + rr;
                if (orow_i >= M) break;
                float4 v = *reinterpret_cast<const float4 *>(stg + rr * lt::STAGE_LD + lane * 4);
                v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                float *o = out + (size_t)orow_i * N + col;
                if (vec_ok && col + 3 < N) *reinterpret_cast<float4 *>(o) = v;
                else {
                    if (col < N) o[0] = v.x;
                    if (col + 1 < N) o[1] = v.y;
                    if (col + 2 < N) o[2] = v.z;
                    if (col + 3 < N) o[3] = v.w;
                }
            }
            __syncwarp();                         // staging tile is rewritten by the next n-tile
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
    }
}

// K (= encode_size) is zero-padded to a multiple of 64 inside the operand images
bool label_tcgen05_shape_ok(const c2v_dims *d) { return d->encode >= 4 && d->encode <= 128 && (d->encode & 3) == 0; }

size_t label_tcgen05_workspace_bytes(const c2v_dims *d, int B)
{
    const size_t nkb = (size_t)(d->encode + 63) / 64;
    const size_t mt = (size_t)(B + 127) / 128, nt = (size_t)(d->label_count + 127) / 128;
    return 1024 + (mt + nt) * nkb * 2 * lt::TILE_BYTES;
}

int launch_loss_argmax(const float *out, const long long *label, int B, long long C, float *loss,
                       long long *argmax, float *maxval, float *d_out, cudaStream_t st);

// argmax / maxval (torch.max(dim=1), main.py:285): a second pass over the logits, which are still in L2
// (B*C*4 = 33.5 MB at cfg2).  Folding it into the GEMM epilogue was measured slower (shuffle argmax 52 us,
// register argmax 44 us, vs 20.6 + 10.3 us for GEMM + this pass; scripts/time_label.py).
int launch_label_tcgen05(const c2v_dims *d, const float *cv, int B, const float *Wout, const float *bias,
                         float *out, long long *argmax, float *maxval, void *ws, size_t ws_bytes, bool reuse_prep,
                         cudaStream_t st)
{
    if (!label_tcgen05_shape_ok(d)) {
        set_error("tcgen05 label GEMM needs encode_size %% 4 == 0 and <= 128 (got %d)", d->encode);
        return C2V_EUNSUPPORTED;
    }
    const int H = d->encode, nkb = (H + 63) / 64;
    const long long C = d->label_count;
    if (!ws || ws_bytes < label_tcgen05_workspace_bytes(d, B)) {
        set_error("label workspace too small: %zu < %zu", ws_bytes, label_tcgen05_workspace_bytes(d, B));
        return C2V_EWORKSPACE;
    }
    uint8_t *p = static_cast<uint8_t *>(ws);
    float *hdr = reinterpret_cast<float *>(p);
    unsigned *mxbits = reinterpret_cast<unsigned *>(p + 256);
    const size_t mt = (size_t)(B + 127) / 128, nt = (size_t)((C + 127) / 128);
    // W_out image first (reusable across calls while the weights are unchanged), cv image, argmax keys
    uint8_t *imgB = p + 1024, *imgA = imgB + nt * nkb * 2 * lt::TILE_BYTES;
    const bool want_arg = argmax || maxval;
    int dev = 0, sms = 0;
    C2V_CUDA_OK(cudaGetDevice(&dev));
    C2V_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));

    if (!reuse_prep) {
        C2V_CUDA_OK(cudaMemsetAsync(mxbits, 0, 4, st));
        absmax_kernel<<<sms * 4, 256, 0, st>>>(Wout, C * H, mxbits);
        C2V_LAUNCH_OK("absmax_kernel");
        split_rows_kernel<<<sms * 8, 256, 0, st>>>(Wout, C, H, nkb, mxbits, imgB, hdr);
        C2V_LAUNCH_OK("split_rows_kernel");
    }
    split_rows_kernel<<<(unsigned)((mt * 128 * nkb * 16 + 255) / 256), 256, 0, st>>>(cv, B, H, nkb, nullptr, imgA, hdr);
    C2V_LAUNCH_OK("split_rows_kernel");

    const int op_bytes = nkb * 2 * lt::TILE_BYTES;
    const int smem_bytes = 2 * op_bytes + 4 * 32 * lt::STAGE_LD * 4 + 128 + 1024;
    C2V_CUDA_OK(cudaFuncSetAttribute(label_gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int G = sms / (int)mt;
    if (G < 1) G = 1;
    if (G > (int)nt) G = (int)nt;
    dim3 grid((unsigned)G, (unsigned)mt);
    label_gemm_tcgen05_kernel<<<grid, 192, smem_bytes, st>>>(imgA, imgB, bias, hdr, out, B, C, nkb, (int)nt);
    C2V_LAUNCH_OK("label_gemm_tcgen05_kernel");
    if (want_arg) return launch_loss_argmax(out, nullptr, B, C, nullptr, argmax, maxval, nullptr, st);
    return C2V_OK;
}

}  // namespace c2v
