// Optimiser step over a flat parameter arena.  This translation unit is compiled WITHOUT --use_fast_math
// (gan_control_b200/build.py NO_FAST_MATH): torch.optim.Adam semantics need IEEE sqrt / division and no
// flush-to-zero of the second moment -- checkpoints interchange with the reference's optimiser state (gt.py:852-865).
#include "common.cuh"

namespace b200gan {

// torch.optim.Adam (gt.py:161-173; eps added after the bias-corrected sqrt, no weight decay, no amsgrad):
//   m = beta1*m + (1-beta1)*g ; v = beta2*v + (1-beta2)*g*g
//   p -= (lr / (1-beta1^t)) * m / (sqrt(v)/sqrt(1-beta2^t) + eps)
// fused with `accumulate` (trainers/utils.py:8-12): ema = ema*decay + p*(1-decay).
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float* ema, float beta1, float beta2,
                                         float eps, float step_size, float rsq_c2, float ema_decay, float grad_scale) {
    const float gi = g * grad_scale;
    m = __fadd_rn(__fmul_rn(beta1, m), __fmul_rn(1.f - beta1, gi));
    v = __fadd_rn(__fmul_rn(beta2, v), __fmul_rn(__fmul_rn(1.f - beta2, gi), gi));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), rsq_c2), eps);
    p = __fsub_rn(p, __fmul_rn(step_size, __fdiv_rn(m, denom)));
    if (ema) *ema = __fadd_rn(__fmul_rn(*ema, ema_decay), __fmul_rn(p, 1.f - ema_decay));
}

template <bool VEC4>
__global__ void __launch_bounds__(256) adam_ema_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v,
                                                       float* __restrict__ ema, int64_t numel, float lr,
                                                       float beta1, float beta2, float eps,
                                                       const float* __restrict__ bias_corr, float ema_decay,
                                                       float grad_scale) {
    // 1 - beta^t lives on the device so that a captured CUDA graph replays with the right step count
    const float step_size = __fdiv_rn(lr, bias_corr[0]), rsq_c2 = __fsqrt_rn(bias_corr[1]);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (VEC4) {
        const int64_t n4 = numel >> 2;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i],
                   vv = reinterpret_cast<float4*>(v)[i];
            const float4 gg = reinterpret_cast<const float4*>(g)[i];
            float4 ee = ema ? reinterpret_cast<float4*>(ema)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            adam_one(pp.x, gg.x, mm.x, vv.x, ema ? &ee.x : nullptr, beta1, beta2, eps, step_size, rsq_c2, ema_decay, grad_scale);
            adam_one(pp.y, gg.y, mm.y, vv.y, ema ? &ee.y : nullptr, beta1, beta2, eps, step_size, rsq_c2, ema_decay, grad_scale);
            adam_one(pp.z, gg.z, mm.z, vv.z, ema ? &ee.z : nullptr, beta1, beta2, eps, step_size, rsq_c2, ema_decay, grad_scale);
            adam_one(pp.w, gg.w, mm.w, vv.w, ema ? &ee.w : nullptr, beta1, beta2, eps, step_size, rsq_c2, ema_decay, grad_scale);
            reinterpret_cast<float4*>(p)[i] = pp;
            reinterpret_cast<float4*>(m)[i] = mm;
            reinterpret_cast<float4*>(v)[i] = vv;
            if (ema) reinterpret_cast<float4*>(ema)[i] = ee;
        }
        for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += stride)
            adam_one(p[i], g[i], m[i], v[i], ema ? ema + i : nullptr, beta1, beta2, eps, step_size, rsq_c2, ema_decay, grad_scale);
    } else {
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < numel; i += stride)
            adam_one(p[i], g[i], m[i], v[i], ema ? ema + i : nullptr, beta1, beta2, eps, step_size, rsq_c2, ema_decay, grad_scale);
    }
}

}  // namespace b200gan

extern "C" int b200gan_adam_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t numel, float lr,
                                float beta1, float beta2, float eps, const float* bias_corr,
                                float ema_decay, float grad_scale, void* stream) {
    using namespace b200gan;
    if (numel <= 0) return 0;
    B200_REQUIRE(bias_corr != nullptr, "adam_ema: bias_corr (device float[2] = {1-beta1^t, 1-beta2^t}) is required");
    const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema) & 15) == 0;
    int64_t blocks = cdiv(vec ? cdiv(numel, 4) : numel, 256);
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (vec)
        adam_ema_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, numel, lr, beta1, beta2,
                                                                                  eps, bias_corr, ema_decay, grad_scale);
    else
        adam_ema_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, ema, numel, lr, beta1, beta2,
                                                                                   eps, bias_corr, ema_decay, grad_scale);
    count_launch();
    return check_launch("adam_ema");
}
