/*
 * c2v_oracle.c -- CPU restatement of the code2vec path-attention forward /
 * backward, used ONLY as test infrastructure (tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg).  It is never linked into, imported by or
 * called from the product path (code2vec_b200/); the product fails loudly when
 * its CUDA library is missing instead of falling back to this file.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors
 * (SURVEY.md section 4), so the pin is made by us: oracle/gen_golden.py imports
 * the unmodified /root/reference/model/model.py, runs it on seeded inputs plus
 * the RNG-free known-answer case of SURVEY.md section 8(c), and commits the
 * inputs/outputs under tests/golden/.  tests/test_oracle_golden.py checks this
 * file against every one of those vectors (<= 2e-6 abs).
 *
 * Arithmetic: every tensor the reference materialises in fp32 is rounded to
 * fp32 here at the same point; reductions (dot products, LayerNorm moments,
 * softmax sums) are accumulated in double and rounded once, i.e. this is the
 * "correctly rounded" version of each ATen op, so the distance to the
 * reference is ATen's own summation-order noise (~1e-7).
 *
 * Reference lines cited as model.py:NN are /root/reference/model/model.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define C2V_NINF (-3.4e38f) /* model.py:12  NINF = -3.4 * 10^38 (finite fp32) */

#ifdef _OPENMP
#include <omp.h>
#endif

int c2v_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* One context row: gathers + concat + input_linear + LayerNorm + tanh (+dropout).
 * model.py:48-51 (three embedding gathers, cat along dim 2: [start ; path ; end]),
 * model.py:54 (Linear, no bias: x = c . W^T, W is [H, D] row-major),
 * model.py:55-56 (LayerNorm over H, biased variance, eps inside sqrt, affine),
 * model.py:57 (tanh), model.py:60-61 (dropout: multiplicative mask already
 * scaled by 1/(1-p); NULL in eval).  Optionally returns xhat and t for backward. */
static void encode_row(const float *es, const float *ep, const float *ee,
                       int Et, int Ep, const float *W, int H,
                       const float *g, const float *b, float eps,
                       const float *drop, float *h, float *xhat_out, float *t_out,
                       float *rstd_out, float *xbuf)
{
    const int D = 2 * Et + Ep;
    for (int o = 0; o < H; ++o) {
        const float *w = W + (size_t)o * D;
        double acc = 0.0;
        for (int k = 0; k < Et; ++k) acc += (double)es[k] * (double)w[k];
        for (int k = 0; k < Ep; ++k) acc += (double)ep[k] * (double)w[Et + k];
        for (int k = 0; k < Et; ++k) acc += (double)ee[k] * (double)w[Et + Ep + k];
        xbuf[o] = (float)acc;
    }
    double mu = 0.0;
    for (int o = 0; o < H; ++o) mu += xbuf[o];
    mu /= H;
    double var = 0.0;
    for (int o = 0; o < H; ++o) { double d = xbuf[o] - mu; var += d * d; }
    var /= H;                                   /* biased, as nn.LayerNorm */
    const double rstd = 1.0 / sqrt(var + (double)eps);
    if (rstd_out) *rstd_out = (float)rstd;
    for (int o = 0; o < H; ++o) {
        const float xh = (float)((xbuf[o] - mu) * rstd);
        const float y = (float)((double)xh * g[o] + b[o]);
        const float t = (float)tanh((double)y);
        if (xhat_out) xhat_out[o] = xh;
        if (t_out) t_out[o] = t;
        h[o] = drop ? t * drop[o] : t;
    }
}

/* Code2Vec.forward up to the code vector: model.py:44-69 and get_attention
 * model.py:90-96.
 *   starts/paths/ends : int64 [B*L]   (main.py:166-168 feed int64 [b, L])
 *   emb_t [T,Et], emb_p [P,Ep], W [H, 2Et+Ep], ln_g/ln_b [H], attn [H]
 *   dropmask : NULL (eval) or [B*L*H] multiplicative mask incl. 1/(1-p)
 *   out: code_vector [B,H], attention [B,L]; optional ctx_h [B*L*H]
 * Mask is (starts > 0) only (model.py:64); masked scores become NINF through
 * score*mask + (1-mask)*NINF (model.py:93) so an all-pad bag is uniform 1/L.
 * Returns 0, or -1 on an out-of-range index (the reference raises IndexError). */
int c2v_oracle_encode_forward(const int64_t *starts, const int64_t *paths, const int64_t *ends,
                              int B, int L,
                              const float *emb_t, int64_t T, int Et,
                              const float *emb_p, int64_t P, int Ep,
                              const float *W, int H,
                              const float *ln_g, const float *ln_b, float ln_eps,
                              const float *attn, const float *dropmask,
                              float *code_vector, float *attention, float *ctx_h)
{
    const int64_t N = (int64_t)B * L;
    for (int64_t i = 0; i < N; ++i) {
        if (starts[i] < 0 || starts[i] >= T || ends[i] < 0 || ends[i] >= T ||
            paths[i] < 0 || paths[i] >= P) return -1;
    }
    int err = 0;
#pragma omp parallel
    {
        float *xbuf = (float *)malloc(sizeof(float) * (size_t)H);
        float *hbag = (float *)malloc(sizeof(float) * (size_t)H * (size_t)L);
        float *z = (float *)malloc(sizeof(float) * (size_t)L);
        if (!xbuf || !hbag || !z) {
#pragma omp atomic write
            err = 1;
        } else {
#pragma omp for schedule(static)
            for (int bag = 0; bag < B; ++bag) {
                for (int j = 0; j < L; ++j) {
                    const int64_t i = (int64_t)bag * L + j;
                    float *h = hbag + (size_t)j * H;
                    encode_row(emb_t + (size_t)starts[i] * Et, emb_p + (size_t)paths[i] * Ep,
                               emb_t + (size_t)ends[i] * Et, Et, Ep, W, H, ln_g, ln_b, ln_eps,
                               dropmask ? dropmask + (size_t)i * H : NULL, h, NULL, NULL, NULL, xbuf);
                    if (ctx_h) memcpy(ctx_h + (size_t)i * H, h, sizeof(float) * (size_t)H);
                    double s = 0.0;                       /* model.py:92-93 score = h . a */
                    for (int o = 0; o < H; ++o) s += (double)h[o] * (double)attn[o];
                    const float m = starts[i] > 0 ? 1.0f : 0.0f;     /* model.py:64 */
                    z[j] = (float)s * m + (1.0f - m) * C2V_NINF;     /* model.py:93 */
                }
                float zmax = z[0];                        /* model.py:96 softmax over dim 1 */
                for (int j = 1; j < L; ++j) if (z[j] > zmax) zmax = z[j];
                double den = 0.0;
                for (int j = 0; j < L; ++j) den += exp((double)z[j] - (double)zmax);
                float *arow = attention + (size_t)bag * L;
                for (int j = 0; j < L; ++j) arow[j] = (float)(exp((double)z[j] - (double)zmax) / den);
                for (int o = 0; o < H; ++o) {             /* model.py:68-69 weighted sum */
                    double acc = 0.0;
                    for (int j = 0; j < L; ++j) acc += (double)arow[j] * (double)hbag[(size_t)j * H + o];
                    code_vector[(size_t)bag * H + o] = (float)acc;
                }
            }
        }
        free(xbuf); free(hbag); free(z);
    }
    return err ? -2 : 0;
}

/* Plain label head: outputs = cv . W_out^T + b   (model.py:83; W_out [C,H]). */
int c2v_oracle_label_logits(const float *cv, int B, int H, const float *Wout, const float *bias,
                            int64_t C, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t idx = 0; idx < (int64_t)B * C; ++idx) {
        const int bag = (int)(idx / C);
        const int64_t c = idx % C;
        const float *v = cv + (size_t)bag * H;
        const float *w = Wout + (size_t)c * H;
        double acc = 0.0;
        for (int o = 0; o < H; ++o) acc += (double)v[o] * (double)w[o];
        out[idx] = (float)(acc + (bias ? (double)bias[c] : 0.0));
    }
    return 0;
}

/* Angular-margin head, model.py:71-80 (uses label in forward, also in eval):
 * cos = normalize(cv) . normalize(W)^T (F.normalize: x / max(||x||, 1e-12));
 * sin = sqrt(1-cos^2); phi = cos*cos(m) - sin*sin(m); phi = cos>0 ? phi : cos;
 * out = (label==c ? phi : cos) * inverse_temp. */
int c2v_oracle_angular_logits(const float *cv, int B, int H, const float *Wout, int64_t C,
                              const int64_t *label, float margin, float inverse_temp, float *out)
{
    const float cos_m = (float)cos((double)margin), sin_m = (float)sin((double)margin);
    float *wn = (float *)malloc(sizeof(float) * (size_t)C);
    if (!wn) return -2;
    for (int64_t c = 0; c < C; ++c) {
        double s = 0.0;
        for (int o = 0; o < H; ++o) s += (double)Wout[c * H + o] * Wout[c * H + o];
        float n = (float)sqrt(s);
        wn[c] = n > 1e-12f ? n : 1e-12f;
    }
    for (int bag = 0; bag < B; ++bag) {
        const float *v = cv + (size_t)bag * H;
        double s = 0.0;
        for (int o = 0; o < H; ++o) s += (double)v[o] * v[o];
        float vn = (float)sqrt(s);
        if (vn < 1e-12f) vn = 1e-12f;
        for (int64_t c = 0; c < C; ++c) {
            double acc = 0.0;
            for (int o = 0; o < H; ++o)
                acc += (double)(v[o] / vn) * (double)(Wout[c * H + o] / wn[c]);
            const float cosv = (float)acc;
            float r = 1.0f - cosv * cosv;
            const float sinv = sqrtf(r);       /* NaN if |cos| rounds above 1, as torch.sqrt */
            float phi = cosv * cos_m - sinv * sin_m;
            if (!(cosv > 0.0f)) phi = cosv;
            out[(size_t)bag * C + c] = (label[bag] == c ? phi : cosv) * inverse_temp;
        }
    }
    free(wn);
    return 0;
}

/* Loss + predict next to the path: main.py:251-264 (log_softmax dim 1, then
 * NLLLoss with weight 1/freq where every freq is 1 => plain mean NLL) and
 * main.py:285 (torch.max(preds, dim=1) -> value, argmax; first max wins). */
int c2v_oracle_loss_argmax(const float *logits, int B, int64_t C, const int64_t *label,
                           float *loss_out, int64_t *argmax_out, float *max_out)
{
    double total = 0.0;
    for (int bag = 0; bag < B; ++bag) {
        const float *r = logits + (size_t)bag * C;
        float mx = r[0]; int64_t am = 0;
        for (int64_t c = 1; c < C; ++c) if (r[c] > mx) { mx = r[c]; am = c; }
        double den = 0.0;
        for (int64_t c = 0; c < C; ++c) den += exp((double)r[c] - (double)mx);
        const double lse = (double)mx + log(den);
        if (label) total += lse - (double)r[label[bag]];
        if (argmax_out) argmax_out[bag] = am;
        if (max_out) max_out[bag] = mx;
    }
    if (loss_out) *loss_out = (float)(total / B);
    return 0;
}

/* Backward of forward + plain head + mean-NLL, hand-derived and
 * pinned against the reference's autograd by tests/golden/grad_*.npz.
 * All math in double; grads are returned as fp32.  dropmask as in forward.
 * g_logits [B,C] is dLoss/doutputs (for mean NLL: (softmax - onehot)/B).
 * Outputs (zero-filled here): d_emb_t [T,Et], d_emb_p [P,Ep], dW [H,D],
 * d_ln_g, d_ln_b, d_attn [H], dWout [C,H], d_bias [C]. */
int c2v_oracle_backward(const int64_t *starts, const int64_t *paths, const int64_t *ends,
                        int B, int L,
                        const float *emb_t, int64_t T, int Et,
                        const float *emb_p, int64_t P, int Ep,
                        const float *W, int H,
                        const float *ln_g, const float *ln_b, float ln_eps,
                        const float *attn, const float *dropmask,
                        const float *Wout, int64_t C, const float *g_logits,
                        float *d_emb_t, float *d_emb_p, float *dW, float *d_ln_g, float *d_ln_b,
                        float *d_attn, float *dWout, float *d_bias)
{
    const int D = 2 * Et + Ep;
    double *gt = (double *)calloc((size_t)T * Et, sizeof(double));
    double *gp = (double *)calloc((size_t)P * Ep, sizeof(double));
    double *gW = (double *)calloc((size_t)H * D, sizeof(double));
    double *gg = (double *)calloc((size_t)H, sizeof(double));
    double *gb = (double *)calloc((size_t)H, sizeof(double));
    double *ga = (double *)calloc((size_t)H, sizeof(double));
    float *xbuf = (float *)malloc(sizeof(float) * (size_t)H);
    float *hb = (float *)malloc(sizeof(float) * (size_t)L * H);
    float *xh = (float *)malloc(sizeof(float) * (size_t)L * H);
    float *tt = (float *)malloc(sizeof(float) * (size_t)L * H);
    float *rs = (float *)malloc(sizeof(float) * (size_t)L);
    double *alpha = (double *)malloc(sizeof(double) * (size_t)L);
    double *zz = (double *)malloc(sizeof(double) * (size_t)L);
    double *cv = (double *)malloc(sizeof(double) * (size_t)H);
    double *gv = (double *)malloc(sizeof(double) * (size_t)H);
    double *dx = (double *)malloc(sizeof(double) * (size_t)H);
    double *dxh = (double *)malloc(sizeof(double) * (size_t)H);
    double *cbuf = (double *)malloc(sizeof(double) * (size_t)D);
    if (!gt || !gp || !gW || !gg || !gb || !ga || !xbuf || !hb || !xh || !tt || !rs || !alpha ||
        !zz || !cv || !gv || !dx || !dxh || !cbuf) return -2;
    memset(dWout, 0, sizeof(float) * (size_t)C * H);
    memset(d_bias, 0, sizeof(float) * (size_t)C);
    double *gWo = (double *)calloc((size_t)C * H, sizeof(double));
    double *gbo = (double *)calloc((size_t)C, sizeof(double));
    if (!gWo || !gbo) return -2;

    for (int bag = 0; bag < B; ++bag) {
        for (int j = 0; j < L; ++j) {
            const int64_t i = (int64_t)bag * L + j;
            encode_row(emb_t + (size_t)starts[i] * Et, emb_p + (size_t)paths[i] * Ep,
                       emb_t + (size_t)ends[i] * Et, Et, Ep, W, H, ln_g, ln_b, ln_eps,
                       dropmask ? dropmask + (size_t)i * H : NULL, hb + (size_t)j * H,
                       xh + (size_t)j * H, tt + (size_t)j * H, rs + j, xbuf);
            double s = 0.0;
            for (int o = 0; o < H; ++o) s += (double)hb[(size_t)j * H + o] * attn[o];
            zz[j] = starts[i] > 0 ? s : (double)C2V_NINF;
        }
        double zmax = zz[0];
        for (int j = 1; j < L; ++j) if (zz[j] > zmax) zmax = zz[j];
        double den = 0.0;
        for (int j = 0; j < L; ++j) { alpha[j] = exp(zz[j] - zmax); den += alpha[j]; }
        for (int j = 0; j < L; ++j) alpha[j] /= den;
        for (int o = 0; o < H; ++o) {
            double acc = 0.0;
            for (int j = 0; j < L; ++j) acc += alpha[j] * hb[(size_t)j * H + o];
            cv[o] = acc;
        }
        /* head: o = Wout v + b ; dWout += g v^T ; db += g ; gv = Wout^T g */
        const float *g = g_logits + (size_t)bag * C;
        for (int o = 0; o < H; ++o) gv[o] = 0.0;
        for (int64_t c = 0; c < C; ++c) {
            const double gc = g[c];
            if (gc == 0.0) continue;
            gbo[c] += gc;
            for (int o = 0; o < H; ++o) {
                gWo[c * H + o] += gc * cv[o];
                gv[o] += gc * (double)Wout[c * H + o];
            }
        }
        double gvv = 0.0;                       /* sum_k alpha_k dalpha_k == gv . v */
        for (int o = 0; o < H; ++o) gvv += gv[o] * cv[o];
        for (int j = 0; j < L; ++j) {
            const int64_t i = (int64_t)bag * L + j;
            const float *h = hb + (size_t)j * H;
            double dalpha = 0.0;
            for (int o = 0; o < H; ++o) dalpha += gv[o] * h[o];
            const double dz = alpha[j] * (dalpha - gvv);
            const double du = starts[i] > 0 ? dz : 0.0;
            double m1 = 0.0, m2 = 0.0;
            for (int o = 0; o < H; ++o) {
                const double dh = alpha[j] * gv[o] + du * attn[o];
                ga[o] += du * h[o];
                const double dmask = dropmask ? dropmask[(size_t)i * H + o] : 1.0;
                const double t = tt[(size_t)j * H + o];
                const double dy = dh * dmask * (1.0 - t * t);
                const double xhv = xh[(size_t)j * H + o];
                gg[o] += dy * xhv;
                gb[o] += dy;
                dxh[o] = dy * ln_g[o];
                m1 += dxh[o];
                m2 += dxh[o] * xhv;
            }
            m1 /= H; m2 /= H;
            int any = 0;
            for (int o = 0; o < H; ++o) {
                dx[o] = rs[j] * (dxh[o] - m1 - (double)xh[(size_t)j * H + o] * m2);
                if (dx[o] != 0.0) any = 1;
            }
            if (!any) continue;
            const float *es = emb_t + (size_t)starts[i] * Et;
            const float *ep = emb_p + (size_t)paths[i] * Ep;
            const float *ee = emb_t + (size_t)ends[i] * Et;
            for (int k = 0; k < Et; ++k) cbuf[k] = es[k];
            for (int k = 0; k < Ep; ++k) cbuf[Et + k] = ep[k];
            for (int k = 0; k < Et; ++k) cbuf[Et + Ep + k] = ee[k];
            double *gs = gt + (size_t)starts[i] * Et;
            double *gpp = gp + (size_t)paths[i] * Ep;
            double *ge = gt + (size_t)ends[i] * Et;
            for (int o = 0; o < H; ++o) {
                const double d = dx[o];
                if (d == 0.0) continue;
                const float *w = W + (size_t)o * D;
                double *gw = gW + (size_t)o * D;
                for (int k = 0; k < D; ++k) gw[k] += d * cbuf[k];
                for (int k = 0; k < Et; ++k) gs[k] += d * w[k];
                for (int k = 0; k < Ep; ++k) gpp[k] += d * w[Et + k];
                for (int k = 0; k < Et; ++k) ge[k] += d * w[Et + Ep + k];
            }
        }
    }
    for (size_t i = 0; i < (size_t)T * Et; ++i) d_emb_t[i] = (float)gt[i];
    for (size_t i = 0; i < (size_t)P * Ep; ++i) d_emb_p[i] = (float)gp[i];
    for (size_t i = 0; i < (size_t)H * D; ++i) dW[i] = (float)gW[i];
    for (int o = 0; o < H; ++o) { d_ln_g[o] = (float)gg[o]; d_ln_b[o] = (float)gb[o]; d_attn[o] = (float)ga[o]; }
    for (size_t i = 0; i < (size_t)C * H; ++i) dWout[i] = (float)gWo[i];
    for (int64_t c = 0; c < C; ++c) d_bias[c] = (float)gbo[c];
    free(gt); free(gp); free(gW); free(gg); free(gb); free(ga); free(xbuf); free(hb); free(xh);
    free(tt); free(rs); free(alpha); free(zz); free(cv); free(gv); free(dx); free(dxh); free(cbuf);
    free(gWo); free(gbo);
    return 0;
}
