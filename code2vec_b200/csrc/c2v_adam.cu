// c2v_adam.cu -- fused flat-buffer Adam: the optimizer side of the training step.
// torch.optim.Adam(model.parameters(), lr, betas, weight_decay) of main.py:138 + optimizer.step() (:175) +
// optimizer.zero_grad() (:171) for ALL parameters in one launch over flat fp32 buffers: reads p, g, m, v once, writes
// p, m, v and the zeroed gradient (ready for the next backward), and folds the 1/world of the data-parallel mean into
// the gradient read (the all_reduce then is a plain sum).  Dense on purpose: momentum keeps moving embedding rows that
// received no gradient, so a row-sparse Adam would not be the reference's optimizer.
// Same operation order as torch's single-tensor Adam (amsgrad=False, maximize=False):
//   g += wd * p;  m += (g - m) * (1 - b1);  v = v * b2 + (1 - b2) * g * g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
#include "c2v_common.cuh"

namespace c2v {

__global__ void __launch_bounds__(256)
adam_step_kernel(float4 *__restrict__ p, float4 *__restrict__ g, float4 *__restrict__ m, float4 *__restrict__ v,
                 long long n4, float *__restrict__ pt, float *__restrict__ gt, float *__restrict__ mt, float *__restrict__ vt,
                 int tail, float step_size, float one_minus_b1, float b2, float one_minus_b2, float inv_sqrt_bc2, float eps,
                 float wd, float gscale, int zero_grad)
{
    auto upd = [&](float &pp, float &gg, float &mm, float &vv) {
        float gr = gg * gscale;
        if (wd != 0.0f) gr = fmaf(wd, pp, gr);
        mm = mm + (gr - mm) * one_minus_b1;                 // exp_avg.lerp_(grad, 1 - beta1)
        vv = vv * b2 + one_minus_b2 * gr * gr;              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
        pp = pp - step_size * (mm / denom);                 // param.addcdiv_(exp_avg, denom, value=-step_size)
        if (zero_grad) gg = 0.0f;
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 P = p[i], G = g[i], M = m[i], V = v[i];
        upd(P.x, G.x, M.x, V.x); upd(P.y, G.y, M.y, V.y); upd(P.z, G.z, M.z, V.z); upd(P.w, G.w, M.w, V.w);
        p[i] = P; m[i] = M; v[i] = V;
        if (zero_grad) g[i] = G;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < tail) {
        const int i = threadIdx.x;
        float P = pt[i], G = gt[i], M = mt[i], V = vt[i];
        upd(P, G, M, V);
        pt[i] = P; mt[i] = M; vt[i] = V;
        if (zero_grad) gt[i] = G;
    }
}

}  // namespace c2v

using namespace c2v;

extern "C" int c2v_adam_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                             int32_t zero_grad, void *stream)
{
    if (!param || !grad || !exp_avg || !exp_avg_sq || n < 1 || step < 1) {
        set_error("c2v_adam_step: bad argument");
        return C2V_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) {
        set_error("c2v_adam_step: buffers must be 16-byte aligned");
        return C2V_EINVAL;
    }
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const long long n4 = n / 4;
    const int tail = (int)(n % 4);
    int dev = 0, sms = 0;
    C2V_CUDA_OK(cudaGetDevice(&dev));
    C2V_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long blocks = (n4 + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    if (blocks < 1) blocks = 1;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4 *>(param), reinterpret_cast<float4 *>(grad), reinterpret_cast<float4 *>(exp_avg),
        reinterpret_cast<float4 *>(exp_avg_sq), n4, param + n4 * 4, grad + n4 * 4, exp_avg + n4 * 4, exp_avg_sq + n4 * 4, tail,
        step_size, 1.0f - beta1, beta2, 1.0f - beta2, inv_sqrt_bc2, eps, weight_decay, grad_scale, zero_grad);
    C2V_LAUNCH_OK("adam_step_kernel");
    return C2V_OK;
}
