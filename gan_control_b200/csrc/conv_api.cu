// extern "C" entry points of the convolution family; picks the engine per call.
#include <atomic>

#include "conv.cuh"

namespace b200gan {
// 0 = automatic (tcgen05 when the shape is eligible), 1 = CUDA-core engine only, 2 = automatic but without
// the halo-reuse variant (testing / A-B timing)
static std::atomic<int> g_conv_engine{0};
}  // namespace b200gan

extern "C" int b200gan_set_conv_engine(int engine) {
    int prev = b200gan::g_conv_engine.exchange(engine & 0xff);
    return prev;
}


extern "C" int b200gan_conv_fwd(const void* x, const void* w, void* y, int dtype, int b, int in_h, int in_w, int ic,
                                int out_h, int out_w, int oc, int kh, int kw, int up, int down, int pad0,
                                int w_per_sample, const float* bias, const float* rowscale, const void* noise,
                                const float* noise_w, float slope, float gain, void* stream) {
    using namespace b200gan;
    ConvGeom g{b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, up, down, pad0, w_per_sample};
    if (g_conv_engine.load() != 1 && b > 0 && b <= 65535 && conv_fwd_pointwise_eligible(dtype, g, x, w, y))
        return conv_fwd_pointwise(x, w, y, dtype, g, bias, rowscale, noise, noise_w, slope, gain, (cudaStream_t)stream);
    if (g_conv_engine.load() == 0 && b > 0 && conv_fwd_halo_eligible(dtype, g, x, w, y))
        return conv_fwd_halo(x, w, y, g, bias, rowscale, noise, noise_w, slope, gain, (cudaStream_t)stream);
    if (g_conv_engine.load() != 1 && b > 0 && conv_fwd_umma_eligible(dtype, g, x, w, y))
        return conv_fwd_umma(x, w, y, g, bias, rowscale, noise, noise_w, slope, gain, (cudaStream_t)stream);
    return conv_fwd_simt(x, w, y, dtype, g, bias, rowscale, noise, noise_w, slope, gain, (cudaStream_t)stream);
}

extern "C" int b200gan_conv_wgrad(const void* x, const void* gy, float* gw, int dtype, int b, int in_h, int in_w,
                                  int ic, int out_h, int out_w, int oc, int kh, int kw, int up, int down, int pad0,
                                  int w_per_sample, void* stream) {
    using namespace b200gan;
    ConvGeom g{b, in_h, in_w, ic, out_h, out_w, oc, kh, kw, up, down, pad0, w_per_sample};
    if (g_conv_engine.load() != 1 && b > 0 && b <= 65535 && conv_wgrad_pointwise_eligible(dtype, g, x, gy))
        return conv_wgrad_pointwise(x, gy, gw, dtype, g, (cudaStream_t)stream);
    if (g_conv_engine.load() == 0 && b > 0 && conv_wgrad_halo_eligible(dtype, g, x, gy))
        return conv_wgrad_halo(x, gy, gw, g, (cudaStream_t)stream);
    if (g_conv_engine.load() != 1 && b > 0 && conv_wgrad_umma_eligible(dtype, g, x, gy))
        return conv_wgrad_umma(x, gy, gw, g, (cudaStream_t)stream);
    return conv_wgrad_simt(x, gy, gw, dtype, g, (cudaStream_t)stream);
}
