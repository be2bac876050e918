// c2v_ffma_tile.cuh -- the 64-row CUDA-core tile GEMM  X = [start ; path ; end] . W^T  shared by the
// FFMA forward kernel (c2v_encode_ffma.cu) and the backward kernel (c2v_backward.cu).
#pragma once
#include "c2v_common.cuh"

namespace c2v {

constexpr int TM = 64;        // context rows per tile
constexpr int KC = 32;        // k-chunk
constexpr int KCP = KC + 4;   // padded A row (keeps 16-B alignment)
constexpr int NB = 128;       // output columns per column block
constexpr int THREADS = 256;

struct FfmaSmem {
    // byte offsets into dynamic smem
    int idx, ac, wc, x, z, e, m, total;
};
__host__ __device__ inline FfmaSmem ffma_smem_layout(int Hs)
{
    FfmaSmem s;
    int o = 0;
    s.idx = o; o += 3 * TM * 8;
    s.ac = o;  o += 2 * TM * KCP * 4;
    s.wc = o;  o += 2 * KC * NB * 4;
    s.x = o;   o += TM * Hs * 4;
    s.z = o;   o += TM * 4;
    s.e = o;   o += TM * 4;
    s.m = o;   o += TM * 4;
    s.total = o;
    return s;
}

template <bool VEC>
__device__ __forceinline__ void load_a_chunk(const EncodeArgs &a, const long long *sidx, float *Ac,
                                             int kc)
{
    // TM rows x KC floats of the concatenated context vector [start ; path ; end] (model.py:51)
    const int D = a.D, Et = a.Et, Ep = a.Ep;
    if (VEC) {
        for (int i = threadIdx.x; i < TM * (KC / 4); i += THREADS) {
            const int r = i / (KC / 4), kq = i % (KC / 4);
            const int k = kc * KC + kq * 4;
            float *dst = Ac + r * KCP + kq * 4;
            if (k < D) {
                const float *src;
                if (k < Et) src = a.emb_t + (size_t)sidx[r] * Et + k;
                else if (k < Et + Ep) src = a.emb_p + (size_t)sidx[TM + r] * Ep + (k - Et);
                else src = a.emb_t + (size_t)sidx[2 * TM + r] * Et + (k - Et - Ep);
                cp_async16(dst, src);
            } else {
                *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    } else {
        for (int i = threadIdx.x; i < TM * KC; i += THREADS) {
            const int r = i / KC, kk = i % KC;
            const int k = kc * KC + kk;
            float *dst = Ac + r * KCP + kk;
            if (k < D) {
                const float *src;
                if (k < Et) src = a.emb_t + (size_t)sidx[r] * Et + k;
                else if (k < Et + Ep) src = a.emb_p + (size_t)sidx[TM + r] * Ep + (k - Et);
                else src = a.emb_t + (size_t)sidx[2 * TM + r] * Et + (k - Et - Ep);
                cp_async4(dst, src);
            } else {
                *dst = 0.0f;
            }
        }
    }
}

__device__ __forceinline__ void load_w_chunk(const float *__restrict__ Wt, int D, int Hs, float *Wc,
                                             int kc, int cb)
{
    for (int i = threadIdx.x; i < KC * (NB / 4); i += THREADS) {
        const int kk = i / (NB / 4), cq = i % (NB / 4);
        const int k = kc * KC + kk, col = cb * NB + cq * 4;
        float *dst = Wc + kk * NB + cq * 4;
        if (k < D && col < Hs) cp_async16(dst, Wt + (size_t)k * Hs + col);
        else *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}


// One 64-row tile: gathers the context rows chunk by chunk (cp.async double buffering) and leaves
// X[r][0..H) = c_r . W^T in shared memory (model.py:48-54).  Needs sidx filled and a __syncthreads
// before; ends with all threads past a __syncthreads.
template <bool VEC>
__device__ __forceinline__ void tile_gemm_xw(const EncodeArgs &a, const long long *sidx, float *Ac,
                                             float *Wc, float *X, int Hs)
{
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int H = a.H, D = a.D;
    const int n_kc = (D + KC - 1) / KC;
    const int n_cb = (H + NB - 1) / NB;
    for (int cb = 0; cb < n_cb; ++cb) {
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

        load_a_chunk<VEC>(a, sidx, Ac, 0);
        load_w_chunk(a.ws.w_t, D, Hs, Wc, 0, cb);
        cp_async_commit();
        for (int kc = 0; kc < n_kc; ++kc) {
            const int buf = kc & 1;
            if (kc + 1 < n_kc) {
                load_a_chunk<VEC>(a, sidx, Ac + (buf ^ 1) * TM * KCP, kc + 1);
                load_w_chunk(a.ws.w_t, D, Hs, Wc + (buf ^ 1) * KC * NB, kc + 1, cb);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const float *Ab = Ac + buf * TM * KCP + (ty * 4) * KCP;
            const float *Wb = Wc + buf * KC * NB;
#pragma unroll
            for (int k4 = 0; k4 < KC; k4 += 4) {
                float4 av[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4 *>(Ab + i * KCP + k4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const float4 w0 = *reinterpret_cast<const float4 *>(Wb + (k4 + kk) * NB + tx * 4);
                    const float4 w1 = *reinterpret_cast<const float4 *>(Wb + (k4 + kk) * NB + 64 + tx * 4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float ai = kk == 0 ? av[i].x : kk == 1 ? av[i].y : kk == 2 ? av[i].z : av[i].w;
                        acc[i][0] = fmaf(ai, w0.x, acc[i][0]); acc[i][1] = fmaf(ai, w0.y, acc[i][1]);
                        acc[i][2] = fmaf(ai, w0.z, acc[i][2]); acc[i][3] = fmaf(ai, w0.w, acc[i][3]);
                        acc[i][4] = fmaf(ai, w1.x, acc[i][4]); acc[i][5] = fmaf(ai, w1.y, acc[i][5]);
                        acc[i][6] = fmaf(ai, w1.z, acc[i][6]); acc[i][7] = fmaf(ai, w1.w, acc[i][7]);
                    }
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = ty * 4 + i;
            const int c0 = cb * NB + tx * 4, c1 = c0 + 64;
            if (c0 < Hs) *reinterpret_cast<float4 *>(X + r * Hs + c0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
            if (c1 < Hs) *reinterpret_cast<float4 *>(X + r * Hs + c1) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
    }
    __syncthreads();
}

// Loads the 3 x 64 indices of a tile into smem; out-of-range ones are clamped to row 0 and counted.
__device__ __forceinline__ void tile_load_indices(const EncodeArgs &a, long long row0, long long *sidx)
{
    for (int i = threadIdx.x; i < 3 * TM; i += THREADS) {
        const int which = i / TM, r = i % TM;
        const long long row = row0 + r;
        long long v = 0;
        if (row < a.N) {
            const long long *src = which == 0 ? a.starts : which == 1 ? a.paths : a.ends;
            v = src[row];
            const long long lim = which == 1 ? a.P : a.T;
            if (v < 0 || v >= lim) { if (a.ws.status) atomicAdd((unsigned long long *)a.ws.status, 1ull); v = 0; }
        }
        sidx[i] = v;
    }
}

}  // namespace c2v
